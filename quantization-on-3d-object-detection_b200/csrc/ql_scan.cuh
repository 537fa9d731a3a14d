// Block-level exclusive scan and the single-CTA scan of per-block counts used by the first-touch numbering
// passes of voxelize.cu and rulebook.cu.
#pragma once
#include "ql_common.cuh"

#define QL_SCAN_THREADS 256

namespace {
constexpr int kScanThreads = QL_SCAN_THREADS;

__device__ __forceinline__ int block_exclusive_scan(int v, int& total) {
    __shared__ int warp_sums[kScanThreads / 32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    int base = wid ? warp_sums[wid - 1] : 0;
    total = warp_sums[kScanThreads / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// single CTA: exclusive scan of block_counts[0..nb) in place, total -> *n_total, min(total, cap) -> *n_out.
// 8 consecutive counts per thread per pass (2048 per pass): the strided rulebook of the finest stage scans ~6 k counts.
__global__ void k_scan_blocks(int* block_counts, int nb, int* n_total, int* n_out, int64_t cap) {
    constexpr int kPer = 8;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += kScanThreads * kPer) {
        const int i0 = base + threadIdx.x * kPer;
        int v[kPer];
        int sum = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            v[j] = i0 + j < nb ? block_counts[i0 + j] : 0;
            sum += v[j];
        }
        int total;
        int run = block_exclusive_scan(sum, total) + carry;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            if (i0 + j < nb) block_counts[i0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (n_total) *n_total = carry;
        if (n_out) *n_out = (int)((int64_t)carry < cap ? (int64_t)carry : cap);
    }
}

}  // namespace
