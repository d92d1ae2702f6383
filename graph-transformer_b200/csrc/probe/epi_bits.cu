// Host evaluation of the bit arithmetic of the FFN chunk epilogue (csrc/ffn_epi.cuh) for the CPU test suite: the same inline
// functions the kernels compile, called on the host.
#include "../ffn_epi.cuh"

extern "C" int u2gnn_epi_bits_host(uint32_t keep, uint32_t* kp16, int* pos32, int* elem32) {
    uint32_t kp[16];
    epi::keep_factors16(keep, kp);
    for (int j = 0; j < 16; ++j) kp16[j] = kp[j];
    for (int e = 0; e < 32; ++e) {
        pos32[e] = epi::flag_pos(e);
        elem32[e] = epi::flag_elem(e);
    }
    return 0;
}
