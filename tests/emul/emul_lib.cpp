// tests/emul/emul_lib.cpp -- TEST INFRASTRUCTURE ONLY: the product's api.cu compiled for the CPU
// SIMT interpreter (cuda_emul.h).  Built by tests/emul/build.sh into tests/emul/libb200lap_emul.so.
#include "cuda_emul.h"
#include "../../gnn-accelerated-lap-warm-start-pipeline_b200/csrc/api.cu"
#include "../../gnn-accelerated-lap-warm-start-pipeline_b200/csrc/host_narrow.cpp"
