// index_storage.h — grow-only device buffers and the owner of one cloud's spatial index.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "internal.h"

namespace se3 {

struct DeviceBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    DeviceBuf() = default;
    DeviceBuf(const DeviceBuf&) = delete;
    DeviceBuf& operator=(const DeviceBuf&) = delete;
    ~DeviceBuf() { release(); }
    // grow-only; contents are NOT preserved across a growth
    int ensure(size_t bytes);
    // grow-only and preserving the first `keep` bytes (used by append)
    int ensure_keep(size_t bytes, size_t keep, cudaStream_t st);
    void release();
    void swap(DeviceBuf& o) {
        void* tp = ptr; ptr = o.ptr; o.ptr = tp;
        size_t tc = cap; cap = o.cap; o.cap = tc;
    }
    template <class T>
    T* as() const {
        return reinterpret_cast<T*>(ptr);
    }
};

struct IndexStorage {
    CloudIndex view;
    DeviceBuf x, y, z, sx, sy, sz, perm, keys, keys_tmp, vals_tmp, box, bbox, bbox_part, sort_tmp;
    size_t sort_tmp_bytes = 0;
    void swap(IndexStorage& o) {
        CloudIndex tv = view; view = o.view; o.view = tv;
        DeviceBuf* a[] = {&x, &y, &z, &sx, &sy, &sz, &perm, &keys, &keys_tmp, &vals_tmp, &box, &bbox, &bbox_part, &sort_tmp};
        DeviceBuf* b[] = {&o.x, &o.y, &o.z, &o.sx, &o.sy, &o.sz, &o.perm, &o.keys, &o.keys_tmp, &o.vals_tmp, &o.box, &o.bbox, &o.bbox_part, &o.sort_tmp};
        for (int i = 0; i < 14; i++) a[i]->swap(*b[i]);
        size_t t = sort_tmp_bytes; sort_tmp_bytes = o.sort_tmp_bytes; o.sort_tmp_bytes = t;
    }
    void plan_levels(int n);
    int reserve(int n);
    int build(cudaStream_t st, long long* launches);
};

// 12-D search structure over the target's SE(3) rows (se3_index.cu)
struct Se3IndexStorage {
    DeviceBuf perm12, inv12, keys12, keys_tmp, vals_tmp, box12, rows32, rows64, sort_tmp;
    size_t sort_tmp_bytes = 0;
    int reserve(int n, const CloudIndex& levels);
    int build(const CloudIndex& I, const double* frame, double alpha, double tscale, IterState* state, cudaStream_t st,
              long long* launches);
};

}  // namespace se3
