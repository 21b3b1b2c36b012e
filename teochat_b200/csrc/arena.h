// Bump allocator over a caller-owned workspace (the C-ABI never allocates): 256-byte aligned slices, and a dry-run mode
// (base == nullptr) that only adds up the size — the `*_workspace_bytes` queries run the same layout code.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace teo {

struct Arena {
    uint8_t* base;
    size_t size, off = 0;
    bool ok = true;
    Arena(void* p, size_t n) : base(static_cast<uint8_t*>(p)), size(n) {}
    template <typename T>
    T* take(size_t count) {
        const size_t bytes = (count * sizeof(T) + 255) & ~static_cast<size_t>(255);
        if (base == nullptr || off + bytes > size) {
            ok = false;
            off += bytes;
            return nullptr;
        }
        T* p = reinterpret_cast<T*>(base + off);
        off += bytes;
        return p;
    }
};
static inline size_t al256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }


}  // namespace teo
