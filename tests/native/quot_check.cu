// quot_check.cu — QuotBy<kCorrected> (rdis_b200/csrc/factors.cuh: one reciprocal, then q = a*y, r = a - b*q, q + r*y) against the
// division instruction sequence on 2^28 operand pairs: the production BA gradient replaces the reference's 24 quotients per
// observation by three reciprocals + corrections and must not change a bit.  Operands: (a) uniform mantissas over a wide
// exponent range, (b) the magnitudes of a bundle-adjustment graph (denominators P_z, P_z^2 in [0.1, 1e4], |r| in [1e-3, 4]).
// Built by rdis_b200/host/Makefile with nvcc for sm_100a; run by tests/test_gpu_parity.py::test_corrected_quotients_are_exact.
#include <cstdio>
#include <cstdlib>

#include "../../rdis_b200/csrc/factors.cuh"

using namespace rdisgpu;

__device__ __forceinline__ unsigned long long splitmix(unsigned long long& s) {
  s += 0x9E3779B97F4A7C15ULL;
  unsigned long long z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

__global__ void check(long long per_thread, unsigned long long* bad, unsigned long long* done) {
  unsigned long long s = 0x1234567ULL + 7919ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x);
  unsigned long long nbad = 0;
  for (long long i = 0; i < per_thread; ++i) {
    const unsigned long long u = splitmix(s), v = splitmix(s), w = splitmix(s);
    double a, b;
    if (i & 1) {  // wide: random mantissas, exponents in [-60, 60]
      a = __longlong_as_double((long long)((u & 0x800FFFFFFFFFFFFFULL) | ((unsigned long long)(1023 - 60 + (w % 121)) << 52)));
      b = __longlong_as_double((long long)((v & 0x800FFFFFFFFFFFFFULL) | ((unsigned long long)(1023 - 60 + ((w >> 20) % 121)) << 52)));
    } else {      // bundle-adjustment magnitudes
      a = ((double)(u >> 11) / 9007199254740992.0 * 2.0 - 1.0) * 1e3;
      b = 0.1 + (double)(v >> 11) / 9007199254740992.0 * ((w & 1) ? 4.0 : 1e4);
      if (w & 2) b = -b;
    }
    const QuotBy<kCorrected> by(b);
    const double q1 = by(a), q0 = a / b;
    if (__double_as_longlong(q0) != __double_as_longlong(q1)) ++nbad;
  }
  atomicAdd(bad, nbad);
  atomicAdd(done, (unsigned long long)per_thread);
}

int main() {
  unsigned long long *bad, *done, hb = 0, hd = 0;
  if (cudaMalloc(&bad, 8) != cudaSuccess || cudaMalloc(&done, 8) != cudaSuccess) {
    std::printf("cuda error\n");
    return 2;
  }
  cudaMemset(bad, 0, 8);
  cudaMemset(done, 0, 8);
  check<<<1024, 256>>>(1024, bad, done);  // 2^28 pairs
  if (cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(&hd, done, 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
    std::printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 2;
  }
  std::printf("checked %llu quotients, mismatches %llu\n", hd, hb);
  return hb == 0 ? 0 : 1;
}
