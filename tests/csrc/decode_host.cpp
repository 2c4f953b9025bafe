// Host emulation of the device-side Uncalled4 BAM decoder (nanocompore_b200/csrc/decode.cu) for the CPU
// test-suite (tests/test_decode_host.py): the same staging code (decode_stage.h) and the same per-unit
// functions the kernels call (inflate_core.cuh, decode_core.cuh), run by loops instead of a grid.  What
// it produces -- the CSR batch of a list of transcripts -- is compared with the host reader + packer.
// Test infrastructure only; built together with reader.cu compiled as C++.
#include <string.h>
#include <vector>

#include "decode_stage.h"
#include "inflate_core.cuh"

using namespace ncb_decode;

extern "C" {

// Returns the number of entries (pos_offsets[P]), -1 when the plan fails, -2 when the outputs are too small.
long long ncb_test_decode(int n_files, ncb_bam *const *bams, const uint8_t *file_label, int n_tx, const int32_t *refs,
                          const int64_t *ref_len, int kit_len, int kit_center, int motor_offset, double max_ratio,
                          int64_t *pos_offsets, int16_t *signal, uint8_t *label, long long cap_entries,
                          int32_t *tx_flags, int64_t *tx_kept) {
    Plan plan;
    if (!make_plan(plan, n_files, bams, n_tx, refs, ref_len)) return -1;
    std::vector<uint32_t> words((size_t)plan.n_words + 2, 0xdeadbeefu);
    stage_payloads(plan, words.data(), 0, plan.blocks.size());
    std::vector<uint8_t> text((size_t)plan.text_bytes + 16, 0);
    std::vector<int32_t> flags((size_t)n_tx, 0);
    static ncb_inflate::Tables tables;
    ncb_inflate::SerialWarp w;
    for (const BlockDesc &b : plan.blocks) {
        const int rc = ncb_inflate::inflate_member(w, tables, words.data() + b.word_off, b.n_words, (int64_t)b.cbytes,
                                                   text.data() + b.dst_off, (int64_t)b.usize);
        if (rc != 0) flags[b.tx] |= TX_ERROR;
    }
    std::vector<ReadRec> recs((size_t)plan.n_slots + 1);
    memset(recs.data(), 0, recs.size() * sizeof(ReadRec));
    for (const SpanDesc &sp : plan.spans)
        if (!flags[sp.tx]) flags[sp.tx] |= walk_span(text.data(), sp, recs.data());
    std::vector<int16_t> rows_i((size_t)plan.row_elems + 1), rows_d((size_t)plan.row_elems + 1);
    Segments sg;
    for (int t = 0; t < n_tx; ++t) {
        const TxDesc &tx = plan.txs[t];
        for (int s = 0; s < tx.n_slots && !flags[t]; ++s) {
            ReadRec &r = recs[tx.first_slot + s];
            if (r.state != SLOT_PRIMARY) continue;
            const int64_t row = tx.row_base + (int64_t)s * tx.ref_len;
            flags[t] |= decode_read(w, text.data(), r, tx.ref_len, kit_len, kit_center, max_ratio, rows_i.data() + row,
                                    rows_d.data() + row, sg);
        }
    }
    std::vector<int32_t> order((size_t)plan.n_slots + 1, -1), n_kept((size_t)n_tx, 0);
    for (int t = 0; t < n_tx; ++t) {
        const TxDesc &tx = plan.txs[t];
        if (flags[t]) continue;
        for (int i = tx.first_slot; i < tx.first_slot + tx.n_slots; ++i) {
            if (recs[i].state != SLOT_KEPT) continue;
            int rank = 0;
            for (int j = tx.first_slot; j < tx.first_slot + tx.n_slots; ++j)
                if (j != i && recs[j].state == SLOT_KEPT && read_before(text.data(), recs[j], j, recs[i], i)) ++rank;
            order[tx.first_slot + rank] = i;
            ++n_kept[t];
        }
    }
    int64_t at = 0;
    for (int t = 0; t < n_tx; ++t) {
        const TxDesc &tx = plan.txs[t];
        for (int p = 0; p < tx.ref_len; ++p) {
            pos_offsets[tx.pos_base + p] = at;
            at += position_count(rows_i.data(), rows_d.data(), tx, order.data(), n_kept[t], p, motor_offset);
        }
    }
    pos_offsets[plan.n_positions] = at;
    if (at > cap_entries) return -2;
    for (int t = 0; t < n_tx; ++t) {
        const TxDesc &tx = plan.txs[t];
        for (int p = 0; p < tx.ref_len; ++p)
            position_fill(rows_i.data(), rows_d.data(), tx, order.data(), n_kept[t], p, motor_offset, recs.data(), file_label,
                          pos_offsets[tx.pos_base + p], signal, label);
        tx_flags[t] = flags[t];
        tx_kept[t] = n_kept[t];
    }
    return at;
}

}
