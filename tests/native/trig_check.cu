// trig_check.cu — demands that rdis_sin / rdis_cos / rdis_sincos (rdis_b200/csrc/factors.cuh) are
// bit-identical to the CUDA math library's sin() / cos() on the device, over a dense random sample of
// the range the factor graphs use, wide ranges, and the special values.  Built by the host Makefile
// with nvcc for sm_100a; run by tests/test_gpu_parity.py::test_inline_trig_is_bit_identical.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../../rdis_b200/csrc/factors.cuh"

using namespace rdisgpu;

__global__ void check(const double* x, long long n, unsigned long long* bad) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = x[i];
    const double s0 = sin(v), c0 = cos(v);
    const double s1 = rdis_sin(v), c1 = rdis_cos(v);
    double s2, c2;
    rdis_sincos(v, s2, c2);
    const bool ok = __double_as_longlong(s0) == __double_as_longlong(s1) && __double_as_longlong(c0) == __double_as_longlong(c1) &&
                    __double_as_longlong(s0) == __double_as_longlong(s2) && __double_as_longlong(c0) == __double_as_longlong(c2);
    const bool both_nan = (s0 != s0) && (s1 != s1) && (s2 != s2) && (c0 != c0) && (c1 != c1) && (c2 != c2);
    if (!ok && !both_nan) atomicAdd(bad, 1ULL);
  }
}

int main() {
  std::vector<double> h;
  unsigned long long st = 88172645463325252ULL;
  auto rnd = [&]() {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    return (double)(st >> 11) / 9007199254740992.0;
  };
  for (int i = 0; i < 4000000; ++i) h.push_back((rnd() * 2 - 1) * 70.0);            // the sinusoid graphs' domain
  for (int i = 0; i < 2000000; ++i) h.push_back((rnd() * 2 - 1) * 1e5);
  for (int i = 0; i < 1000000; ++i) h.push_back((rnd() * 2 - 1) * 4e9);             // crosses the 2^31 slow-path threshold
  for (int i = 0; i < 1000000; ++i) h.push_back(std::ldexp(rnd() * 2 - 1, (int)(rnd() * 600) - 300));
  for (int k = -40; k <= 40; ++k)                                                      // around multiples of pi/2
    for (int d = -3; d <= 3; ++d) h.push_back(std::nextafter(k * 1.5707963267948966, d > 0 ? 1e300 : -1e300) + d * 1e-17);
  const double sp[] = {0.0, -0.0, INFINITY, -INFINITY, NAN, 2147483648.0, -2147483648.0, 2147483647.9999998, 1e308, 5e-324};
  for (double v : sp) h.push_back(v);
  double* d;
  unsigned long long* bad;
  if (cudaMalloc(&d, h.size() * sizeof(double)) != cudaSuccess || cudaMalloc(&bad, 8) != cudaSuccess) {
    std::printf("cuda error\n");
    return 2;
  }
  cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
  cudaMemset(bad, 0, 8);
  check<<<592, 256>>>(d, (long long)h.size(), bad);
  unsigned long long hb = 0;
  if (cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
    std::printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 2;
  }
  std::printf("checked %zu values, mismatches %llu\n", h.size(), hb);
  return hb == 0 ? 0 : 1;
}
