// Host-side management of the paged KV cache (SURVEY.md §8b "teo_kv_alloc/free"): which page ids of the caller-owned
// per-layer pools belong to which sequence.  Replaces the reference's torch.cat growth of past_key_values (HF-4.31
// LlamaAttention, called through llava_llama.py:88-99), where every decoded token re-allocates and copies the whole cache.
// Pure host code: the pools themselves are device memory the caller allocates (teo_kv_pool_bytes per layer); a page id
// indexes the same slot in every layer's pool.
#include <algorithm>
#include <new>
#include <vector>

#include "common.h"

struct teo_kv_allocator {
    int n_pages = 0;
    std::vector<int> free_list;        // kept sorted descending: pages are handed out lowest id first (deterministic layouts)
    std::vector<uint8_t> used;
};

using namespace teo;

extern "C" size_t teo_kv_pool_bytes(int n_pages, int n_heads, int page_size, int head_dim, int exact) {
    if (n_pages <= 0 || n_heads <= 0 || page_size <= 0 || head_dim <= 0) return 0;
    return static_cast<size_t>(n_pages) * 2 * n_heads * page_size * head_dim * (exact ? sizeof(float) : sizeof(bf16));
}

extern "C" int teo_kv_plan(const int* host_seq_lens, int n_seqs, int max_new_tokens, int page_size, int* host_pages_per_seq,
                           int* max_pages_out, int* total_pages_out) {
    TEO_CHECK_ARG(host_seq_lens && n_seqs > 0 && max_new_tokens >= 0 && page_size > 0, "kv_plan: bad arguments");
    long long total = 0;
    int mx = 0;
    for (int i = 0; i < n_seqs; ++i) {
        TEO_CHECK_ARG(host_seq_lens[i] >= 0, "kv_plan: negative length for sequence %d", i);
        const long long need = (static_cast<long long>(host_seq_lens[i]) + max_new_tokens + page_size - 1) / page_size;
        TEO_CHECK_ARG(need < (1 << 30), "kv_plan: sequence %d needs too many pages", i);
        if (host_pages_per_seq) host_pages_per_seq[i] = static_cast<int>(need);
        mx = std::max(mx, static_cast<int>(need));
        total += need;
    }
    TEO_CHECK_ARG(total < (1LL << 31), "kv_plan: %lld pages overflow int", total);
    if (max_pages_out) *max_pages_out = mx;
    if (total_pages_out) *total_pages_out = static_cast<int>(total);
    return TEO_OK;
}

extern "C" int teo_kv_create(int n_pages, teo_kv_allocator** out) {
    TEO_CHECK_ARG(out != nullptr && n_pages > 0, "kv_create: bad arguments");
    teo_kv_allocator* a = new (std::nothrow) teo_kv_allocator();
    TEO_CHECK_ARG(a != nullptr, "kv_create: out of host memory");
    a->n_pages = n_pages;
    a->used.assign(n_pages, 0);
    a->free_list.resize(n_pages);
    for (int i = 0; i < n_pages; ++i) a->free_list[i] = n_pages - 1 - i;
    *out = a;
    return TEO_OK;
}
extern "C" int teo_kv_destroy(teo_kv_allocator* a) {
    delete a;
    return TEO_OK;
}
extern "C" int teo_kv_available(const teo_kv_allocator* a) { return a ? static_cast<int>(a->free_list.size()) : 0; }

/* pages for a sequence that will hold up to n_tokens tokens: writes ceil(n_tokens / page_size) ids (lowest free ids first,
 * ascending) to host_pages_out[0 .. max_pages) and returns their number; TEO_ERR_WORKSPACE when the pool is exhausted
 * (nothing is taken then). */
extern "C" int teo_kv_alloc(teo_kv_allocator* a, int n_tokens, int page_size, int* host_pages_out, int max_pages) {
    TEO_CHECK_ARG(a && host_pages_out && n_tokens > 0 && page_size > 0, "kv_alloc: bad arguments");
    const int need = (n_tokens + page_size - 1) / page_size;
    TEO_CHECK_ARG(need <= max_pages, "kv_alloc: %d pages needed, row holds %d", need, max_pages);
    if (need > static_cast<int>(a->free_list.size())) {
        set_error("kv_alloc: pool exhausted (%d pages needed, %zu free of %d)", need, a->free_list.size(), a->n_pages);
        return TEO_ERR_WORKSPACE;
    }
    for (int i = 0; i < need; ++i) {
        const int p = a->free_list.back();
        a->free_list.pop_back();
        a->used[p] = 1;
        host_pages_out[i] = p;
    }
    return need;
}

extern "C" int teo_kv_free(teo_kv_allocator* a, const int* host_pages, int n_pages) {
    TEO_CHECK_ARG(a && (host_pages || n_pages == 0) && n_pages >= 0, "kv_free: bad arguments");
    for (int i = 0; i < n_pages; ++i) {
        const int p = host_pages[i];
        TEO_CHECK_ARG(p >= 0 && p < a->n_pages && a->used[p], "kv_free: page %d is not allocated", p);
    }
    for (int i = 0; i < n_pages; ++i) {
        a->used[host_pages[i]] = 0;
        a->free_list.push_back(host_pages[i]);
    }
    std::sort(a->free_list.begin(), a->free_list.end(), [](int x, int y) { return x > y; });
    return TEO_OK;
}
