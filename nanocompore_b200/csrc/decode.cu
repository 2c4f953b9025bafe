// Device-side Uncalled4 BAM decoder (the readers -> CSR step, moved onto the GPU).
//
// The host path (reader.cu + pack.cu) inflates the BGZF members of a transcript with zlib, decodes the
// ur / uc / ul tags into planar float rows, packs them into a pinned CSR batch and ncb_compare copies that
// to the device: ~6 ms of host work per 2 000-position transcript at 400 reads, 0.5 ms of GPU work.  Here
// the host only looks the members up in the index of the BAM handles and copies their COMPRESSED bytes into
// pinned memory (decode_stage.h); everything else runs on the device and ends in a device-resident
// NCB_LAYOUT_CSR_I16 batch that ncb_compare takes as it is:
//
//   inflate_kernel   one warp per BGZF member: lane 0 walks the DEFLATE bit stream into a queue of
//                    literal / copy commands, the warp executes them (inflate_core.cuh)
//   walk_kernel      one thread per run of records: primary reads, names, tag arrays
//   read_kernel      one warp per read: ur segments, reversed uc / ul, skips resolved, invalid-kmer ratio
//                    -> int16 row pair
//   order_kernel     rank of every kept read among the reads of its transcript (by name)
//   count_kernel / fill_kernel   one thread per position: entries of the kept reads in that order
//   (CUB exclusive scan between the two: pos_offsets)
//
// Replaces, for a batch of transcripts, pysam fetch + Uncalled4.get_data (uncalled4.py:57-201), the read
// ordering and invalid-kmer filter of Uncalled4Worker._read_data (run.py:668-718) and the tensor assembly.
// The per-unit logic is host/device code (decode_core.cuh) that the CPU tests run against the host path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cub/device/device_scan.cuh>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "decode_stage.h"
#include "inflate_core.cuh"
#include "ncb.h"

using namespace ncb_decode;

namespace {

#define FULL 0xffffffffu

struct CudaWarp {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int stride() const { return 32; }
    __device__ __forceinline__ bool leader() const { return (threadIdx.x & 31) == 0; }
    __device__ __forceinline__ int bcast(int v) const { return __shfl_sync(FULL, v, 0); }
    __device__ __forceinline__ bool any(int v) const { return __any_sync(FULL, v) != 0; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    // bytes of the member's own output written earlier by lanes of this warp (ordered by sync())
    __device__ __forceinline__ uint8_t load_out(const uint8_t *p) const { return *reinterpret_cast<const volatile uint8_t *>(p); }
    __device__ __forceinline__ int64_t reduce_min(int64_t v) const {
        for (int m = 16; m > 0; m >>= 1) { const int64_t o = __shfl_xor_sync(FULL, v, m); v = o < v ? o : v; }
        return v;
    }
    __device__ __forceinline__ int64_t reduce_max(int64_t v) const {
        for (int m = 16; m > 0; m >>= 1) { const int64_t o = __shfl_xor_sync(FULL, v, m); v = o > v ? o : v; }
        return v;
    }
    __device__ __forceinline__ int reduce_add(int v) const {
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
        return v;
    }
};

// A group of G lanes as the "warp" of inflate_member (inflate_core.cuh): 32 / G members per warp, each group on its
// own member, tables and command queue.  The walk of a member's bit stream is one lane's work (91 % of the warp
// instructions the kernel issues have one active lane), so narrower groups looked like a way to share the walk's
// instructions among several members -- measured, they are slower: 0.371 / 0.418 / 0.503 / 0.687 ms per transcript
// at G = 32 / 16 / 8 / 4 (the leaders of a warp do not stay converged through the data-dependent branches of the
// walk, and the cooperative phases get 32 / G times as many trips).  G = 32 is the default; NCB_INFLATE_GROUP
// selects the others for A/B runs.
template <int G>
struct CudaGroup {
    unsigned mask;
    __device__ __forceinline__ CudaGroup() {
        const int l = threadIdx.x & 31;
        mask = G == 32 ? FULL : ((G == 32 ? 0u : (1u << (G & 31)) - 1u) << (l & ~(G - 1)));
    }
    __device__ __forceinline__ int lane() const { return threadIdx.x & (G - 1); }
    __device__ __forceinline__ int stride() const { return G; }
    __device__ __forceinline__ bool leader() const { return (threadIdx.x & (G - 1)) == 0; }
    __device__ __forceinline__ int bcast(int v) const { return __shfl_sync(mask, v, 0, G); }
    __device__ __forceinline__ bool any(int v) const { return __any_sync(mask, v) != 0; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ uint8_t load_out(const uint8_t *p) const { return *reinterpret_cast<const volatile uint8_t *>(p); }
};

constexpr int INFLATE_MEMBERS_PER_CTA = 8;   // 8 x 4.4 KB of tables per CTA, whatever the group width
constexpr int READ_WPC = 8;

template <int G>
__global__ void __launch_bounds__(INFLATE_MEMBERS_PER_CTA * G)
inflate_kernel(const BlockDesc *blocks, int n_blocks, const uint32_t *words, uint8_t *text, int32_t *flags) {
    __shared__ ncb_inflate::Tables tables[INFLATE_MEMBERS_PER_CTA];
    const int slot = threadIdx.x / G;
    const int b = blockIdx.x * INFLATE_MEMBERS_PER_CTA + slot;
    if (b >= n_blocks) return;
    const BlockDesc d = blocks[b];
    CudaGroup<G> w;
    const int rc = ncb_inflate::inflate_member(w, tables[slot], words + d.word_off, d.n_words,
                                               (int64_t)d.cbytes, text + d.dst_off, (int64_t)d.usize);
    if (rc != ncb_inflate::INF_OK && w.leader()) atomicOr(&flags[d.tx], TX_ERROR);
}

__global__ void walk_kernel(const SpanDesc *spans, int n_spans, const uint8_t *text, ReadRec *recs, int32_t *flags) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_spans) return;
    const SpanDesc sp = spans[s];
    if (flags[sp.tx]) return;
    const int st = walk_span(text, sp, recs);
    if (st != TX_OK) atomicOr(&flags[sp.tx], st);
}

__global__ void __launch_bounds__(READ_WPC * 32)
read_kernel(const TxDesc *txs, ReadRec *recs, int n_slots, const uint8_t *text, int16_t *rows_i, int16_t *rows_d,
            int kit_len, int kit_center, double max_invalid_ratio, int32_t *flags) {
    __shared__ Segments segs[READ_WPC];
    const int slot = blockIdx.x * READ_WPC + (threadIdx.x >> 5);
    if (slot >= n_slots) return;
    ReadRec &r = recs[slot];
    if (r.state != SLOT_PRIMARY) return;
    const int t = r.tx;
    if (flags[t]) return;
    const TxDesc tx = txs[t];
    const int64_t row = tx.row_base + (int64_t)(slot - tx.first_slot) * tx.ref_len;
    CudaWarp w;
    const int st = decode_read(w, text, r, (int64_t)tx.ref_len, kit_len, kit_center, max_invalid_ratio, rows_i + row,
                               rows_d + row, segs[threadIdx.x >> 5]);
    if (st != TX_OK && w.leader()) atomicOr(&flags[t], st);
}

__global__ void order_kernel(const TxDesc *txs, const ReadRec *recs, int n_slots, const uint8_t *text,
                             const int32_t *flags, int32_t *order, int32_t *n_kept) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const ReadRec &a = recs[i];
    if (a.state != SLOT_KEPT || flags[a.tx]) return;
    const TxDesc tx = txs[a.tx];
    int rank = 0;
    for (int j = tx.first_slot; j < tx.first_slot + tx.n_slots; ++j)
        if (j != i && recs[j].state == SLOT_KEPT && read_before(text, recs[j], j, a, i)) ++rank;
    order[tx.first_slot + rank] = i;
    atomicAdd(&n_kept[a.tx], 1);
}

// transcript of a global position (binary search over pos_base)
__device__ __forceinline__ int tx_of(const TxDesc *txs, int n_tx, int64_t g) {
    int lo = 0, hi = n_tx - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (txs[mid].pos_base <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void count_kernel(const TxDesc *txs, int n_tx, int64_t n_positions, const int16_t *rows_i, const int16_t *rows_d,
                             const int32_t *order, const int32_t *n_kept, const int32_t *flags, int motor_offset,
                             int64_t *counts) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g > n_positions) return;
    if (g == n_positions) { counts[g] = 0; return; }
    const int t = tx_of(txs, n_tx, g);
    const TxDesc tx = txs[t];
    counts[g] = flags[t] ? 0 : position_count(rows_i, rows_d, tx, order, n_kept[t], (int)(g - tx.pos_base), motor_offset);
}

__global__ void fill_kernel(const TxDesc *txs, int n_tx, int64_t n_positions, const int16_t *rows_i, const int16_t *rows_d,
                            const int32_t *order, const int32_t *n_kept, const int32_t *flags, int motor_offset,
                            const ReadRec *recs, const uint8_t *file_label, const int64_t *pos_offsets, int16_t *signal,
                            uint8_t *label) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_positions) return;
    const int t = tx_of(txs, n_tx, g);
    if (flags[t]) return;
    const TxDesc tx = txs[t];
    position_fill(rows_i, rows_d, tx, order, n_kept[t], (int)(g - tx.pos_base), motor_offset, recs, file_label,
                  pos_offsets[g], signal, label);
}

struct Buf {
    void *ptr = nullptr;
    size_t bytes = 0;
    bool pinned = false;
};

struct Slot {
    Plan plan;
    int32_t n_files = 0, n_tx = 0;
    std::vector<uint8_t> file_label;
    bool staged = false, launched = false;
    // pinned host: payload words, tables (blocks | spans | txs | labels), results (flags | kept | entries)
    Buf h_words, h_tables, h_result;
    // device
    Buf d_words, d_tables, d_text, d_recs, d_rows_i, d_rows_d, d_order, d_state, d_counts, d_offsets, d_signal, d_label,
        d_scan;
    cudaEvent_t done = nullptr;
    int64_t launches = 0;
};

}  // namespace

struct ncb_decoder {
    int device = 0;
    int inflate_group = 32;    // lanes per BGZF member in inflate_kernel (NCB_INFLATE_GROUP = 4 / 8 / 16 / 32: A/B runs)
    std::string err;
    std::vector<Slot> slots;
};

namespace {

int dfail(ncb_decoder *d, int code, const std::string &what, cudaError_t e = cudaSuccess) {
    if (d) {
        d->err = what;
        if (e != cudaSuccess) d->err += std::string(": ") + cudaGetErrorString(e);
    }
    return code;
}

#define DCK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            cudaGetLastError();                                                                \
            return dfail(d, e_ == cudaErrorMemoryAllocation ? NCB_ERR_NOMEM : NCB_ERR_CUDA, #call, e_); \
        }                                                                                      \
    } while (0)

int grow(ncb_decoder *d, Buf &b, size_t bytes, bool pinned) {
    if (bytes == 0) bytes = 256;
    if (b.bytes >= bytes) return NCB_OK;
    if (b.ptr) {
        if (b.pinned) cudaFreeHost(b.ptr); else cudaFree(b.ptr);
        b.ptr = nullptr;
        b.bytes = 0;
    }
    const size_t want = bytes + bytes / 4;
    const cudaError_t e = pinned ? cudaHostAlloc(&b.ptr, want, cudaHostAllocDefault) : cudaMalloc(&b.ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.ptr = nullptr;
        return dfail(d, NCB_ERR_NOMEM, pinned ? "cudaHostAlloc" : "cudaMalloc", e);
    }
    b.bytes = want;
    b.pinned = pinned;
    return NCB_OK;
}
#define GROW(buf, bytes, pinned)                          \
    do {                                                  \
        int r_ = grow(d, (buf), (bytes), (pinned));       \
        if (r_ != NCB_OK) return r_;                      \
    } while (0)

void drop(Buf &b) {
    if (b.ptr) { if (b.pinned) cudaFreeHost(b.ptr); else cudaFree(b.ptr); }
    b.ptr = nullptr;
    b.bytes = 0;
}

struct Guard {      // the caller's current device is put back on return
    int prev = -1;
    bool changed = false;
    cudaError_t enter(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev == device) return cudaSuccess;
        const cudaError_t e = cudaSetDevice(device);
        changed = e == cudaSuccess && prev >= 0;
        return e;
    }
    ~Guard() { if (changed) cudaSetDevice(prev); }
};

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace

extern "C" {

int ncb_decoder_create(ncb_decoder **out, int device, int n_slots) {
    if (!out || n_slots < 1 || n_slots > 16) return NCB_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return NCB_ERR_CUDA;
    }
    ncb_decoder *d = new (std::nothrow) ncb_decoder();
    if (!d) return NCB_ERR_NOMEM;
    d->device = device;
    if (const char *g = getenv("NCB_INFLATE_GROUP")) {
        const int v = atoi(g);
        if (v == 4 || v == 8 || v == 16 || v == 32) d->inflate_group = v;
    }
    try {
        d->slots.resize((size_t)n_slots);
    } catch (const std::exception &) {
        delete d;
        return NCB_ERR_NOMEM;
    }
    Guard g;
    if (g.enter(device) != cudaSuccess) { delete d; return NCB_ERR_CUDA; }
    for (Slot &s : d->slots)
        if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            for (Slot &t : d->slots) if (t.done) cudaEventDestroy(t.done);
            delete d;
            return NCB_ERR_CUDA;
        }
    *out = d;
    return NCB_OK;
}

void ncb_decoder_destroy(ncb_decoder *d) {
    if (!d) return;
    Guard g;
    g.enter(d->device);
    cudaDeviceSynchronize();
    for (Slot &s : d->slots) {
        for (Buf *b : {&s.h_words, &s.h_tables, &s.h_result, &s.d_words, &s.d_tables, &s.d_text, &s.d_recs, &s.d_rows_i,
                       &s.d_rows_d, &s.d_order, &s.d_state, &s.d_counts, &s.d_offsets, &s.d_signal, &s.d_label, &s.d_scan})
            drop(*b);
        if (s.done) cudaEventDestroy(s.done);
    }
    delete d;
}

const char *ncb_decoder_last_error(const ncb_decoder *d) { return d ? d->err.c_str() : "null decoder"; }

int ncb_decoder_stage(ncb_decoder *d, int slot, int32_t n_files, ncb_bam *const *bams, const uint8_t *file_label,
                      int32_t n_tx, const int32_t *refs, const int64_t *ref_len, int threads) {
    if (!d) return NCB_ERR_INVALID;
    if (slot < 0 || slot >= (int)d->slots.size() || n_files < 1 || n_files > 128 || n_tx < 1 || !bams || !file_label ||
        !refs || !ref_len)
        return dfail(d, NCB_ERR_INVALID, "ncb_decoder_stage: bad argument");
    for (int32_t t = 0; t < n_tx; ++t)
        if (ref_len[t] < 0 || ref_len[t] > 0x3fffffff) return dfail(d, NCB_ERR_INVALID, "ncb_decoder_stage: ref_len out of range");
    Slot &s = d->slots[(size_t)slot];
    Guard g;
    DCK(g.enter(d->device));
    if (s.launched) DCK(cudaEventSynchronize(s.done));      // the upload of the slot's last batch reads its pinned buffers
    s.staged = s.launched = false;
    try {
        if (!make_plan(s.plan, n_files, bams, n_tx, refs, ref_len)) return dfail(d, NCB_ERR_INVALID, s.plan.error);
        s.file_label.assign(file_label, file_label + n_files);
    } catch (const std::exception &) {
        return dfail(d, NCB_ERR_NOMEM, "ncb_decoder_stage: out of memory");
    }
    const Plan &p = s.plan;
    if (p.n_slots > 0x7ffffff0LL || p.n_positions > 0x3fffffffLL)
        return dfail(d, NCB_ERR_INVALID, "ncb_decoder_stage: batch too large");
    s.n_files = n_files;
    s.n_tx = n_tx;
    GROW(s.h_words, ((size_t)p.n_words + 2) * sizeof(uint32_t), true);
    const size_t tb = align16(p.blocks.size() * sizeof(BlockDesc)) + align16(p.spans.size() * sizeof(SpanDesc)) +
                      align16(p.txs.size() * sizeof(TxDesc)) + align16((size_t)n_files);
    GROW(s.h_tables, tb, true);
    GROW(s.h_result, align16((size_t)n_tx * 4) + align16((size_t)n_tx * 4) + 16, true);
    uint8_t *t8 = (uint8_t *)s.h_tables.ptr;
    memcpy(t8, p.blocks.data(), p.blocks.size() * sizeof(BlockDesc));
    t8 += align16(p.blocks.size() * sizeof(BlockDesc));
    memcpy(t8, p.spans.data(), p.spans.size() * sizeof(SpanDesc));
    t8 += align16(p.spans.size() * sizeof(SpanDesc));
    memcpy(t8, p.txs.data(), p.txs.size() * sizeof(TxDesc));
    t8 += align16(p.txs.size() * sizeof(TxDesc));
    memcpy(t8, s.file_label.data(), (size_t)n_files);
    // payload copies: page-cache reads of the mapped files, shared out over a few threads
    uint32_t *words = (uint32_t *)s.h_words.ptr;
    const size_t nb = p.blocks.size();
    int nt = threads < 1 ? 1 : (threads > 16 ? 16 : threads);
    if ((size_t)nt > nb / 8 + 1) nt = (int)(nb / 8 + 1);
    if (nt <= 1) {
        stage_payloads(p, words, 0, nb);
    } else {
        std::vector<std::thread> pool;
        const size_t chunk = (nb + nt - 1) / nt;
        for (int k = 0; k < nt; ++k) {
            const size_t lo = k * chunk, hi = lo + chunk < nb ? lo + chunk : nb;
            if (lo >= hi) continue;
            try {
                pool.emplace_back([&p, words, lo, hi] { stage_payloads(p, words, lo, hi); });
            } catch (const std::exception &) {
                stage_payloads(p, words, lo, hi);
            }
        }
        for (auto &th : pool) th.join();
    }
    s.staged = true;
    return NCB_OK;
}

int ncb_decoder_launch(ncb_decoder *d, int slot, int32_t kit_len, int32_t kit_center, int32_t motor_offset,
                       double max_invalid_ratio, ncb_batch *batch, void *stream) {
    if (!d) return NCB_ERR_INVALID;
    if (slot < 0 || slot >= (int)d->slots.size() || !batch || motor_offset < 0 || kit_len < 1)
        return dfail(d, NCB_ERR_INVALID, "ncb_decoder_launch: bad argument");
    Slot &s = d->slots[(size_t)slot];
    if (!s.staged) return dfail(d, NCB_ERR_INVALID, "ncb_decoder_launch: nothing staged in this slot");
    cudaStream_t st = (cudaStream_t)stream;
    Guard g;
    DCK(g.enter(d->device));
    const Plan &p = s.plan;
    const int n_tx = s.n_tx, V = motor_offset > 0 ? 3 : 2;
    const size_t nblocks = p.blocks.size(), nspans = p.spans.size();
    const size_t b_blocks = align16(nblocks * sizeof(BlockDesc)), b_spans = align16(nspans * sizeof(SpanDesc)),
                 b_txs = align16((size_t)n_tx * sizeof(TxDesc)), b_lab = align16((size_t)s.n_files);
    GROW(s.d_words, ((size_t)p.n_words + 2) * sizeof(uint32_t), false);
    GROW(s.d_tables, b_blocks + b_spans + b_txs + b_lab, false);
    GROW(s.d_text, (size_t)p.text_bytes + 64, false);
    GROW(s.d_recs, ((size_t)p.n_slots + 1) * sizeof(ReadRec), false);
    GROW(s.d_rows_i, ((size_t)p.row_elems + 8) * sizeof(int16_t), false);
    GROW(s.d_rows_d, ((size_t)p.row_elems + 8) * sizeof(int16_t), false);
    GROW(s.d_order, ((size_t)p.n_slots + 1) * sizeof(int32_t), false);
    GROW(s.d_state, 2 * align16((size_t)n_tx * 4), false);
    GROW(s.d_counts, ((size_t)p.n_positions + 1) * sizeof(int64_t), false);
    GROW(s.d_offsets, ((size_t)p.n_positions + 1) * sizeof(int64_t), false);
    GROW(s.d_signal, ((size_t)p.row_elems + 8) * V * sizeof(int16_t), false);
    GROW(s.d_label, (size_t)p.row_elems + 8, false);
    size_t scan_bytes = 0;
    DCK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int64_t *)nullptr, (int64_t *)nullptr,
                                      (int)(p.n_positions + 1), st));
    GROW(s.d_scan, scan_bytes, false);

    uint8_t *tab = (uint8_t *)s.d_tables.ptr;
    const BlockDesc *blocks = (const BlockDesc *)tab;
    const SpanDesc *spans = (const SpanDesc *)(tab + b_blocks);
    const TxDesc *txs = (const TxDesc *)(tab + b_blocks + b_spans);
    const uint8_t *labels = tab + b_blocks + b_spans + b_txs;
    int32_t *flags = (int32_t *)s.d_state.ptr;
    int32_t *n_kept = (int32_t *)((uint8_t *)s.d_state.ptr + align16((size_t)n_tx * 4));
    ReadRec *recs = (ReadRec *)s.d_recs.ptr;
    uint8_t *text = (uint8_t *)s.d_text.ptr;
    int16_t *rows_i = (int16_t *)s.d_rows_i.ptr, *rows_d = (int16_t *)s.d_rows_d.ptr;
    int32_t *order = (int32_t *)s.d_order.ptr;
    int64_t *counts = (int64_t *)s.d_counts.ptr, *offsets = (int64_t *)s.d_offsets.ptr;

    DCK(cudaMemcpyAsync(s.d_words.ptr, s.h_words.ptr, ((size_t)p.n_words + 2) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    DCK(cudaMemcpyAsync(s.d_tables.ptr, s.h_tables.ptr, b_blocks + b_spans + b_txs + b_lab, cudaMemcpyHostToDevice, st));
    DCK(cudaMemsetAsync(s.d_state.ptr, 0, 2 * align16((size_t)n_tx * 4), st));
    DCK(cudaMemsetAsync(s.d_recs.ptr, 0, ((size_t)p.n_slots + 1) * sizeof(ReadRec), st));
    s.launches = 0;
    if (nblocks) {
        const unsigned grid = (unsigned)((nblocks + INFLATE_MEMBERS_PER_CTA - 1) / INFLATE_MEMBERS_PER_CTA);
        const uint32_t *words = (const uint32_t *)s.d_words.ptr;
        switch (d->inflate_group) {
            case 16: inflate_kernel<16><<<grid, INFLATE_MEMBERS_PER_CTA * 16, 0, st>>>(blocks, (int)nblocks, words, text, flags); break;
            case 4: inflate_kernel<4><<<grid, INFLATE_MEMBERS_PER_CTA * 4, 0, st>>>(blocks, (int)nblocks, words, text, flags); break;
            case 8: inflate_kernel<8><<<grid, INFLATE_MEMBERS_PER_CTA * 8, 0, st>>>(blocks, (int)nblocks, words, text, flags); break;
            default: inflate_kernel<32><<<grid, INFLATE_MEMBERS_PER_CTA * 32, 0, st>>>(blocks, (int)nblocks, words, text, flags); break;
        }
        ++s.launches;
    }
    if (nspans) {
        walk_kernel<<<(unsigned)((nspans + 63) / 64), 64, 0, st>>>(spans, (int)nspans, text, recs, flags);
        ++s.launches;
    }
    if (p.n_slots) {
        read_kernel<<<(unsigned)((p.n_slots + READ_WPC - 1) / READ_WPC), READ_WPC * 32, 0, st>>>(
            txs, recs, (int)p.n_slots, text, rows_i, rows_d, kit_len, kit_center, max_invalid_ratio, flags);
        order_kernel<<<(unsigned)((p.n_slots + 127) / 128), 128, 0, st>>>(txs, recs, (int)p.n_slots, text, flags, order, n_kept);
        s.launches += 2;
    }
    count_kernel<<<(unsigned)((p.n_positions + 1 + 127) / 128), 128, 0, st>>>(txs, n_tx, p.n_positions, rows_i, rows_d, order,
                                                                               n_kept, flags, motor_offset, counts);
    DCK(cub::DeviceScan::ExclusiveSum(s.d_scan.ptr, scan_bytes, (const int64_t *)counts, offsets, (int)(p.n_positions + 1), st));
    if (p.n_positions) {
        fill_kernel<<<(unsigned)((p.n_positions + 127) / 128), 128, 0, st>>>(
            txs, n_tx, p.n_positions, rows_i, rows_d, order, n_kept, flags, motor_offset, recs, labels, offsets,
            (int16_t *)s.d_signal.ptr, (uint8_t *)s.d_label.ptr);
        ++s.launches;
    }
    s.launches += 2;
    DCK(cudaGetLastError());
    // flags | kept | entries -> pinned host
    uint8_t *hr = (uint8_t *)s.h_result.ptr;
    DCK(cudaMemcpyAsync(hr, flags, (size_t)n_tx * 4, cudaMemcpyDeviceToHost, st));
    DCK(cudaMemcpyAsync(hr + align16((size_t)n_tx * 4), n_kept, (size_t)n_tx * 4, cudaMemcpyDeviceToHost, st));
    DCK(cudaMemcpyAsync(hr + 2 * align16((size_t)n_tx * 4), offsets + p.n_positions, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DCK(cudaEventRecord(s.done, st));
    s.launched = true;

    memset(batch, 0, sizeof(*batch));
    batch->layout = NCB_LAYOUT_CSR_I16;
    batch->memory = NCB_MEM_DEVICE;
    batch->n_positions = p.n_positions;
    batch->n_entries = p.row_elems;            // upper bound: every read at every position
    batch->pos_offsets = offsets;
    batch->signal = s.d_signal.ptr;
    batch->label = (const uint8_t *)s.d_label.ptr;
    batch->n_reads = p.max_slots_per_tx > 0 ? p.max_slots_per_tx : 1;
    batch->n_vars = V;
    return NCB_OK;
}

int ncb_decoder_collect(ncb_decoder *d, int slot, int32_t *tx_flags, int32_t *tx_reads, int64_t *n_entries) {
    if (!d) return NCB_ERR_INVALID;
    if (slot < 0 || slot >= (int)d->slots.size()) return dfail(d, NCB_ERR_INVALID, "ncb_decoder_collect: bad slot");
    Slot &s = d->slots[(size_t)slot];
    if (!s.launched) return dfail(d, NCB_ERR_INVALID, "ncb_decoder_collect: nothing launched in this slot");
    Guard g;
    DCK(g.enter(d->device));
    DCK(cudaEventSynchronize(s.done));
    const uint8_t *hr = (const uint8_t *)s.h_result.ptr;
    if (tx_flags) memcpy(tx_flags, hr, (size_t)s.n_tx * 4);
    if (tx_reads) memcpy(tx_reads, hr + align16((size_t)s.n_tx * 4), (size_t)s.n_tx * 4);
    if (n_entries) memcpy(n_entries, hr + 2 * align16((size_t)s.n_tx * 4), sizeof(int64_t));
    return NCB_OK;
}

int ncb_decoder_download(ncb_decoder *d, int slot, int64_t *pos_offsets, int16_t *signal, uint8_t *label,
                         int64_t cap_entries, int32_t n_vars) {
    if (!d) return NCB_ERR_INVALID;
    if (slot < 0 || slot >= (int)d->slots.size() || !pos_offsets || (n_vars != 2 && n_vars != 3))
        return dfail(d, NCB_ERR_INVALID, "ncb_decoder_download: bad argument");
    Slot &s = d->slots[(size_t)slot];
    if (!s.launched) return dfail(d, NCB_ERR_INVALID, "ncb_decoder_download: nothing launched in this slot");
    Guard g;
    DCK(g.enter(d->device));
    DCK(cudaEventSynchronize(s.done));
    const int64_t P = s.plan.n_positions;
    DCK(cudaMemcpy(pos_offsets, s.d_offsets.ptr, (size_t)(P + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    const int64_t E = pos_offsets[P];
    if (E > cap_entries) return dfail(d, NCB_ERR_INVALID, "ncb_decoder_download: more entries than the caller's capacity");
    if (E > 0 && signal) DCK(cudaMemcpy(signal, s.d_signal.ptr, (size_t)E * n_vars * sizeof(int16_t), cudaMemcpyDeviceToHost));
    if (E > 0 && label) DCK(cudaMemcpy(label, s.d_label.ptr, (size_t)E, cudaMemcpyDeviceToHost));
    return NCB_OK;
}

int64_t ncb_decoder_staged_bytes(const ncb_decoder *d, int slot, int64_t *text_bytes, int64_t *n_members, int64_t *n_records) {
    if (!d || slot < 0 || slot >= (int)d->slots.size() || !d->slots[(size_t)slot].staged) return -1;
    const Plan &p = d->slots[(size_t)slot].plan;
    if (text_bytes) *text_bytes = p.text_bytes;
    if (n_members) *n_members = (int64_t)p.blocks.size();
    if (n_records) *n_records = p.n_slots;
    return p.n_words * 4;
}

int64_t ncb_decoder_kernel_launches(const ncb_decoder *d, int slot) {
    return (d && slot >= 0 && slot < (int)d->slots.size()) ? d->slots[(size_t)slot].launches : 0;
}

}  // extern "C"
