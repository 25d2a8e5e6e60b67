// components.cuh — sibling components on the device.
//
// The reference keeps the connected components of the bipartite variable / factor graph incrementally
// (Holm–de Lichtenberg–Thorup levels over Euler-tour splay trees, src/ConnectivityGraph.cpp:82-520) and reads them
// back as Component::createChildren (src/Component.cpp:508-549): an edge (v, f) exists iff v is unassigned and f is
// not an assigned constant (ConnectivityGraph.cpp:211-285).  That structure is inherently sequential; what the tree
// search consumes is only the MEMBERSHIP.  Here the labels are recomputed from scratch by min-label propagation with
// pointer jumping over the factor CSR — integer work, atomicMin only, so the result is exact and order-independent:
//   var_label[v] = smallest variable id of v's component (-1 for an assigned variable)
//   fac_label[f] = that label for the component factor f belongs to (-1: assigned constant, or no unassigned variable)
#pragma once
#include "factors.cuh"

namespace rdisgpu {

struct ComponentsView {
  const uint8_t* assigned;  // u8[V]
  int32_t* vlabel;          // i32[V]
  int32_t* flabel;          // i32[F]
  int32_t* changed;         // one flag
};

__global__ void cc_init_kernel(GraphView G, ComponentsView C) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < G.V; v += (int64_t)gridDim.x * blockDim.x)
    C.vlabel[v] = C.assigned[v] ? -1 : (int32_t)v;
}

// slot -> variable id of factor f
__device__ __forceinline__ int cc_arity(const GraphView& G, int64_t f) {
  return (G.kind == KIND_NLPF) ? (__ldg(&G.rowptr[f + 1]) - __ldg(&G.rowptr[f])) : 12;
}
__device__ __forceinline__ int32_t cc_var(const GraphView& G, int64_t f, int s) {
  if (G.kind == KIND_NLPF) return __ldg(&G.evid[__ldg(&G.rowptr[f]) + s]);
  return BaOps::slot_vid(G, __ldg(&G.cam[f]), __ldg(&G.pt[f]), s);
}

// One hooking round: every live factor takes the minimum label of its unassigned variables and pushes it back to
// them (atomicMin: the minimum does not depend on the order of arrival).
__global__ void cc_hook_kernel(GraphView G, ComponentsView C) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < G.F; f += (int64_t)gridDim.x * blockDim.x) {
    if (G.fconst_on != nullptr && G.fconst_on[f]) {
      C.flabel[f] = -1;
      continue;
    }
    const int ar = cc_arity(G, f);
    int32_t m = 0x7fffffff;
    for (int s = 0; s < ar; ++s) {
      const int32_t l = C.vlabel[cc_var(G, f, s)];
      if (l >= 0 && l < m) m = l;
    }
    if (m == 0x7fffffff) {
      C.flabel[f] = -1;
      continue;
    }
    C.flabel[f] = m;
    for (int s = 0; s < ar; ++s) {
      const int32_t v = cc_var(G, f, s);
      if (C.vlabel[v] > m) {
        atomicMin(&C.vlabel[v], m);
        *C.changed = 1;
      }
    }
  }
}

// Pointer jumping: a label is the id of a variable of the same component, whose own label is at most as large.
__global__ void cc_jump_kernel(GraphView G, ComponentsView C) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < G.V; v += (int64_t)gridDim.x * blockDim.x) {
    int32_t l = C.vlabel[v];
    if (l < 0) continue;
    int32_t ll = C.vlabel[l];
    while (ll < l) {  // follow the chain to its current root
      l = ll;
      ll = C.vlabel[l];
    }
    if (l < C.vlabel[v]) {
      C.vlabel[v] = l;  // only this thread writes vlabel[v] in this kernel; readers see either value, both valid
      *C.changed = 1;
    }
  }
}

// Final pass: factor labels from the converged variable labels.
__global__ void cc_factor_labels_kernel(GraphView G, ComponentsView C) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < G.F; f += (int64_t)gridDim.x * blockDim.x) {
    int32_t m = -1;
    if (!(G.fconst_on != nullptr && G.fconst_on[f])) {
      const int ar = cc_arity(G, f);
      for (int s = 0; s < ar; ++s) {
        const int32_t l = C.vlabel[cc_var(G, f, s)];
        if (l >= 0) {
          m = l;
          break;  // all unassigned variables of one factor carry the same converged label
        }
      }
    }
    C.flabel[f] = m;
  }
}

}  // namespace rdisgpu
