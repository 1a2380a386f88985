// emu_capi.cpp — TEST INFRASTRUCTURE: rsq_capi.cu (the product's C-ABI layer) compiled for the host over the emulated runtime of
// cuda_emu.h; with emu_kernels.cpp and emu_covstats.cpp this is oracle/_build/librsq_emu.so: the WHOLE product library running on
// the CPU (kernel sources unchanged), so that the -m gpu parity tests can be replayed without a GPU (tests/test_emu_library.py).
#define RSQ_CUDA_EMU 1
#include "cuda_emu.h"

#include "../../rnaseqc_b200/csrc/rsq_capi.cu"
