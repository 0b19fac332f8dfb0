// Target of the link-time aliases for the Marlin kernel instantiations that oracle/ref/Makefile does
// not build (gptq_gemm.cu takes the address of every instantiation its generator would emit).
// Reaching it means a test selected a configuration outside the built set.
#include <cstdio>
#include <cstdlib>

extern "C" void b200_ref_missing_kernel() {
  std::fprintf(stderr, "oracle/_ref: this Marlin kernel instantiation was not built (oracle/ref/Makefile)\n");
  std::abort();
}
