// Host half of the device-side BAM decoder: from the index of the BAM handles (reader.cu) to the tables
// and the staged compressed payloads the kernels of decode.cu work on.  Plain C++ (no CUDA), shared by
// decode.cu and the host emulation of the CPU tests (tests/csrc/decode_host.cpp).
#pragma once
#include <string>
#include <vector>

#include "bam_internal.h"
#include "decode_core.cuh"

namespace ncb_decode {

struct Plan {
    std::vector<BlockDesc> blocks;
    std::vector<SpanDesc> spans;
    std::vector<TxDesc> txs;
    std::vector<const uint8_t *> payload;   // per block: its DEFLATE payload in the file mapping
    int64_t n_words = 0;                    // staged payload words (padding included)
    int64_t text_bytes = 0;                 // inflated bytes of all spans
    int64_t n_slots = 0;                    // read slots (records) of all transcripts
    int64_t row_elems = 0;                  // int16 elements of one row plane (sum of slots x ref_len)
    int64_t n_positions = 0;
    int32_t max_slots_per_tx = 0;
    std::string error;
};

// Walks the runs of every (transcript, file) and lists their BGZF members.  refs[t * n_files + f] is the
// reference index of transcript t in file f (< 0: the file has no reads of it).  Returns false with
// plan.error set when a member header does not parse.
inline bool make_plan(Plan &plan, int32_t n_files, ncb_bam *const *bams, int32_t n_tx, const int32_t *refs,
                      const int64_t *ref_len) {
    plan = Plan();
    for (int32_t t = 0; t < n_tx; ++t) {
        TxDesc tx;
        tx.row_base = plan.row_elems;
        tx.pos_base = plan.n_positions;
        tx.ref_len = (int32_t)ref_len[t];
        tx.first_slot = (int32_t)plan.n_slots;
        tx.n_slots = 0;
        tx.pad = 0;
        for (int32_t f = 0; f < n_files; ++f) {
            const int32_t ref = refs[(size_t)t * n_files + f];
            ncb_bam *b = bams[f];
            if (ref < 0 || !b || ref >= (int32_t)b->runs.size()) continue;
            for (const Run &run : b->runs[ref]) {
                SpanDesc sp;
                sp.text_off = plan.text_bytes;
                sp.text_len = 0;
                sp.in_block = run.in_block;
                sp.n_records = (int32_t)run.n_records;
                sp.tx = t;
                sp.file = f;
                sp.first_slot = (int32_t)(plan.n_slots + tx.n_slots);
                sp.pad = 0;
                for (int64_t off = run.block_off;;) {
                    Block bl;
                    if (!b->block_at(off, bl)) {
                        plan.error = b->path + ": " + (b->error.empty() ? "unexpected end of file" : b->error.c_str());
                        return false;
                    }
                    const uint8_t *src = b->map + bl.file_off;
                    const int hdr = 12 + (int)rd16(src + 10);
                    if (bl.usize > 0) {
                        BlockDesc d;
                        d.cbytes = (uint32_t)(bl.csize - hdr - 8);
                        d.n_words = (d.cbytes + 3) / 4;
                        d.word_off = (uint64_t)plan.n_words;
                        d.usize = (uint32_t)bl.usize;
                        d.tx = t;
                        d.dst_off = plan.text_bytes + sp.text_len;
                        plan.blocks.push_back(d);
                        plan.payload.push_back(src + hdr);
                        plan.n_words += (int64_t)d.n_words + 2;       // two zero words behind every payload
                        sp.text_len += bl.usize;
                    }
                    if (off == run.end_block_off) break;
                    off += bl.csize;
                }
                plan.text_bytes += (sp.text_len + 7) & ~(int64_t)7;
                tx.n_slots += sp.n_records;
                plan.spans.push_back(sp);
            }
        }
        plan.n_slots += tx.n_slots;
        plan.row_elems += (int64_t)tx.n_slots * tx.ref_len;
        plan.n_positions += tx.ref_len;
        if (tx.n_slots > plan.max_slots_per_tx) plan.max_slots_per_tx = tx.n_slots;
        plan.txs.push_back(tx);
    }
    return true;
}

// Copies the payloads of blocks [lo, hi) into the staging words (zero padded to whole words + 2 words).
inline void stage_payloads(const Plan &plan, uint32_t *words, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
        const BlockDesc &d = plan.blocks[i];
        uint32_t *dst = words + d.word_off;
        dst[d.n_words ? d.n_words - 1 : 0] = 0;
        dst[d.n_words] = 0;
        dst[d.n_words + 1] = 0;
        memcpy(dst, plan.payload[i], d.cbytes);
    }
}

}  // namespace ncb_decode
