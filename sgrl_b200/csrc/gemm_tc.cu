// Translation unit(s) of the tcgen05 GEMM kernels: compiled once per part (-DSGRL_TC_PART=0..4, see Makefile and gemm_tc.cuh);
// gemm_tc_api.h says why they are compiled on their own.
#define SGRL_TC_TU 1
#include "gemm_tc.cuh"
