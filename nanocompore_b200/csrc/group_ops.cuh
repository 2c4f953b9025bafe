// Cooperative primitives of one "group" = the threads that share one position:
// a warp (GROUP == 32; several groups per CTA, synchronised with __syncwarp) or a whole
// CTA (GROUP == 256 / 1024; __syncthreads).  Reductions return the SAME bits to every
// thread of the group (xor butterflies are commutative), so branches on reduced values
// are uniform.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ncb {

typedef unsigned long long u64;

#define NCB_FULL 0xffffffffu
#define NCB_RED_MAX 13   /* most 64-bit values reduced in one call */
#define NCB_RED_STRIDE 13 /* row stride of a warp group's reduction scratch (odd: no bank conflicts);
                             warp groups reduce at most 13 values at once */
#define NCB_RED_WARP_WORDS (32 * NCB_RED_STRIDE + 16)   /* 64-bit words of scratch per warp group */

struct OpAddF64 { typedef double T; __device__ static T id() { return 0.0; }
                  __device__ static T f(T a, T b) { return a + b; } };
struct OpMinF64 { typedef double T; __device__ static T id() { return __longlong_as_double(0x7ff0000000000000LL); }
                  __device__ static T f(T a, T b) { return fmin(a, b); } };
struct OpMaxF64 { typedef double T; __device__ static T id() { return __longlong_as_double(0xfff0000000000000LL); }
                  __device__ static T f(T a, T b) { return fmax(a, b); } };
struct OpAddU64 { typedef u64 T; __device__ static T id() { return 0ull; }
                  __device__ static T f(T a, T b) { return a + b; } };
struct OpMaxU64 { typedef u64 T; __device__ static T id() { return 0ull; }
                  __device__ static T f(T a, T b) { return a > b ? a : b; } };
struct OpMinU64 { typedef u64 T; __device__ static T id() { return ~0ull; }
                  __device__ static T f(T a, T b) { return a < b ? a : b; } };

__device__ __forceinline__ double shfl_xor_t(double v, int m) { return __shfl_xor_sync(NCB_FULL, v, m); }
__device__ __forceinline__ u64 shfl_xor_t(u64 v, int m) { return __shfl_xor_sync(NCB_FULL, v, m); }
__device__ __forceinline__ double shfl_up_t(double v, int d) { return __shfl_up_sync(NCB_FULL, v, d); }
__device__ __forceinline__ u64 shfl_up_t(u64 v, int d) { return __shfl_up_sync(NCB_FULL, v, d); }
__device__ __forceinline__ double shfl_idx_t(double v, int l) { return __shfl_sync(NCB_FULL, v, l); }
__device__ __forceinline__ u64 shfl_idx_t(u64 v, int l) { return __shfl_sync(NCB_FULL, v, l); }

template <int GROUP>
struct Grp {
    static constexpr int NW = GROUP / 32;
    u64 *red;     // CTA groups: [2][32][NCB_RED_MAX] scratch; warp groups: [NCB_RED_WARP_WORDS]
    int phase;
    int tid;      // thread index inside the group
    int lane, warp;

    __device__ Grp(u64 *red_, int tid_) : red(red_), phase(0), tid(tid_) {
        lane = tid_ & 31;
        warp = tid_ >> 5;
    }

    __device__ __forceinline__ void sync() const {
        if (GROUP == 32) __syncwarp(); else __syncthreads();
    }

    // all-reduce NV values with Op; result in v[] on every thread.
    // Warp groups with more than two values go through the group's shared-memory scratch
    // (red: [32][NCB_RED_STRIDE] partials + [16] results): every lane stores its NV partials,
    // lane l folds column l % 16 over 16 rows in a fixed order, the two half-warps are combined
    // with one shuffle and the NV results are read back by all lanes -- about 2 NV + 45
    // instructions where a butterfly of 64-bit shuffles costs about 30 NV.
    template <class Op, int NV>
    __device__ __forceinline__ void all_reduce(typename Op::T (&v)[NV]) {
        typedef typename Op::T T;
        static_assert(NV <= NCB_RED_MAX, "too many values in one reduction");
        if (GROUP == 32 && NV > 2) {
            T *buf = reinterpret_cast<T *>(red);
            T *res = buf + 32 * NCB_RED_STRIDE;
#pragma unroll
            for (int i = 0; i < NV; ++i) buf[lane * NCB_RED_STRIDE + i] = v[i];
            __syncwarp();
            const int c = lane & 15;
            const T *col = buf + (lane & 16) * NCB_RED_STRIDE + c;
            T acc = Op::id();
            if (c < NV) {
                T a0 = Op::f(Op::f(col[0 * NCB_RED_STRIDE], col[1 * NCB_RED_STRIDE]),
                             Op::f(col[2 * NCB_RED_STRIDE], col[3 * NCB_RED_STRIDE]));
                T a1 = Op::f(Op::f(col[4 * NCB_RED_STRIDE], col[5 * NCB_RED_STRIDE]),
                             Op::f(col[6 * NCB_RED_STRIDE], col[7 * NCB_RED_STRIDE]));
                T a2 = Op::f(Op::f(col[8 * NCB_RED_STRIDE], col[9 * NCB_RED_STRIDE]),
                             Op::f(col[10 * NCB_RED_STRIDE], col[11 * NCB_RED_STRIDE]));
                T a3 = Op::f(Op::f(col[12 * NCB_RED_STRIDE], col[13 * NCB_RED_STRIDE]),
                             Op::f(col[14 * NCB_RED_STRIDE], col[15 * NCB_RED_STRIDE]));
                acc = Op::f(Op::f(a0, a1), Op::f(a2, a3));
            }
            acc = Op::f(acc, shfl_xor_t(acc, 16));
            if (lane < NV) res[lane] = acc;
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = res[i];
            __syncwarp();
            return;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) v[i] = Op::f(v[i], shfl_xor_t(v[i], m));
        }
        if (GROUP > 32) {
            T *buf = reinterpret_cast<T *>(red) + phase * (32 * NCB_RED_MAX);
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i) buf[warp * NCB_RED_MAX + i] = v[i];
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                T x = buf[(lane % NW) * NCB_RED_MAX + i];
#pragma unroll
                for (int m = NW / 2; m > 0; m >>= 1) x = Op::f(x, shfl_xor_t(x, m));
                v[i] = x;
            }
            phase ^= 1;
        }
    }

    template <class Op>
    __device__ __forceinline__ typename Op::T all_reduce1(typename Op::T x) {
        typename Op::T v[1] = {x};
        all_reduce<Op, 1>(v);
        return v[0];
    }

    // exclusive scan over the threads of the group (thread order), returns the prefix of
    // this thread; *total (optional) receives the reduction over the whole group
    template <class Op>
    __device__ __forceinline__ typename Op::T scan_exclusive(typename Op::T x,
                                                             typename Op::T *total = nullptr) {
        typedef typename Op::T T;
        T inc = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T y = shfl_up_t(inc, d);
            if (lane >= d) inc = Op::f(y, inc);
        }
        T excl = shfl_up_t(inc, 1);
        if (lane == 0) excl = Op::id();
        if (GROUP == 32) {
            if (total) *total = shfl_idx_t(inc, 31);
            return excl;
        }
        T *buf = reinterpret_cast<T *>(red) + phase * (32 * NCB_RED_MAX);
        if (lane == 31) buf[warp] = inc;
        __syncthreads();
        T w = (lane < NW) ? buf[lane] : Op::id();
        T winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T y = shfl_up_t(winc, d);
            if (lane >= d) winc = Op::f(y, winc);
        }
        T wexcl = shfl_up_t(winc, 1);
        if (lane == 0) wexcl = Op::id();
        T mine = shfl_idx_t(wexcl, warp);
        if (total) *total = shfl_idx_t(winc, 31);
        phase ^= 1;
        return Op::f(mine, excl);
    }
};

}  // namespace ncb
