// C-ABI shim around the UNMODIFIED reference sampler class, compiled from the reference's own
// sources where they lie (see oracle/Makefile).  Output goes to oracle/_ref/ only.  Test
// infrastructure: used to pin oracle/log_uniform_oracle.c and the CUDA sampler.
#include "Log_Uniform_Sampler.h"
#include <cstdint>
#include <vector>

extern "C" {
void* logu_ref_new(int n) { return new Log_Uniform_Sampler(n); }
void logu_ref_free(void* s) { delete static_cast<Log_Uniform_Sampler*>(s); }
float logu_ref_probability(void* s, int idx) { return static_cast<Log_Uniform_Sampler*>(s)->probability(idx); }
// ids come back in the reference's own (libstdc++ bucket) order
int logu_ref_sample(void* s, int64_t size, int64_t* out_ids, int* num_tries) {
    auto set = static_cast<Log_Uniform_Sampler*>(s)->sample(static_cast<size_t>(size), num_tries);
    int64_t i = 0;
    for (long v : set) out_ids[i++] = v;
    return 0;
}
void logu_ref_expected_count(void* s, int tries, const int64_t* ids, int64_t n, float* out) {
    std::vector<long> v(ids, ids + n);
    auto f = static_cast<Log_Uniform_Sampler*>(s)->expected_count(tries, v);
    for (int64_t i = 0; i < n; ++i) out[i] = f[i];
}
int logu_ref_sample_unique(void* s, int64_t size, const int64_t* labels, int64_t n_labels, int64_t* out_ids) {
    std::unordered_set<long> l(labels, labels + n_labels);
    auto set = static_cast<Log_Uniform_Sampler*>(s)->sample_unique(static_cast<size_t>(size), l);
    int64_t i = 0;
    for (long v : set) out_ids[i++] = v;
    return 0;
}
}
