// CPU check of the forward-mode arithmetic: the per-cell / per-boundary-entry functions of
// hydrograd.jl_b200/csrc/hg_jvp_impl.h -- the same source the kernels of hg_jvp.cu are built from -- run in plain loops.
// Test infrastructure (built by tests/test_jvp_cpu.py into oracle/_build); not part of the product library.
#include <vector>

#include "../hydrograd.jl_b200/csrc/hg_jvp_impl.h"

using namespace hg::jvp;

template <class T>
static void run(const Args& a) {
  std::vector<T> coef((size_t)a.n_inlet);
  for (int32_t k = 0; k < a.n_inlet; ++k) coef[k] = inlet_coef<T>(a, k);
  for (int32_t e = 0; e < a.B; ++e)
    ghost_entry<T>(a, e, a.bc_type[e] == kInletQ ? coef[a.bc_group[e]] : lift<T>(0.0, 0.0));
  for (int32_t i = 0; i < a.N; ++i) cell<T>(a, i);
}

extern "C" int jvp_host(const Args* a, int dual) {
  if (dual) run<Dual>(*a);
  else run<double>(*a);
  return *a->err;
}
