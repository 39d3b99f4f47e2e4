// TEST INFRASTRUCTURE — not product code.
// Probes the rehash schedule of the C++ standard library this oracle is built against, empirically:
// it inserts distinct keys into a live std::unordered_map<size_t,int> (the container the reference uses,
// grid_subsampling.cpp:46) and records every (size before insert, new bucket_count) at which the table
// rehashes. oracle/kp_oracle.c's unordered_map order model consumes the schedule.
#include <unordered_map>
#include <cstdint>
#include <cstddef>

extern "C" int umap_rehash_schedule(int64_t max_n, int64_t* elt, int64_t* bkt, int cap) {
    std::unordered_map<size_t, int> m;
    size_t bc = m.bucket_count();
    int n = 0;
    for (int64_t i = 0; i < max_n; i++) {
        m.emplace((size_t)i, 0);
        if (m.bucket_count() != bc) {
            bc = m.bucket_count();
            if (n < cap) { elt[n] = i; bkt[n] = (int64_t)bc; }
            n++;
        }
    }
    return n;
}

// Iteration order of a live unordered_map after inserting keys[0..n) (distinct) — used by the tests to pin
// the C model in kp_oracle.c against the real container.
extern "C" void umap_live_order(const uint64_t* keys, int n, int* order) {
    std::unordered_map<size_t, int> m;
    for (int i = 0; i < n; i++) m.emplace((size_t)keys[i], i);
    int r = 0;
    for (auto& v : m) order[r++] = v.second;
}
