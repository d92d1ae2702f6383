// Library-level entry points: error strings, version, device check, host view of the dropout stream.
#include "common.cuh"
#include "rng.cuh"

extern "C" const char* u2gnn_strerror(int code) {
    switch (code) {
        case U2GNN_OK: return "ok";
        case U2GNN_EINVAL: return "invalid argument (shape, null pointer or range)";
        case U2GNN_EALIGN: return "pointer or leading dimension not aligned";
        case U2GNN_EUNSUPPORTED: return "size outside the supported range of the sm_100a kernels";
        case U2GNN_ELAUNCH: return "CUDA launch error";
        case U2GNN_EWORKSPACE: return "workspace too small";
        case U2GNN_EDEVICE: return "no sm_100 device";
        default: return "unknown u2gnn error";
    }
}

extern "C" int u2gnn_version(void) { return 100; }

extern "C" int u2gnn_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return U2GNN_EDEVICE;
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
        (void)cudaGetLastError();
        return U2GNN_EDEVICE;
    }
    return U2GNN_OK;
}

extern "C" uint32_t u2gnn_rng_mask_word_host(uint64_t seed, uint32_t stream, uint64_t group, int thr) {
    if (thr <= 0) return 0xFFFFFFFFu;
    if (thr > 255) return 0u;
    return rng_keep_word(rng_keys(seed, stream), group, thr);
}
