// Device-side Uncalled4 BAM decoder, the per-unit logic as host/device functions (kernels: decode.cu;
// host emulation for the CPU tests: tests/csrc/decode_host.cpp).
//
// From the inflated text of a transcript's BGZF members to the CSR batch ncb_compare takes
// (NCB_LAYOUT_CSR_I16), without the planar float rows and the pinned CSR copy of the host path:
//
//   walk_span      the records of one run of one file: primary reads, their names and the ur / uc / ul
//                  arrays                                    (reader.cu: ncb_bam_read_uncalled4, find_signal_tags)
//   decode_read    one read -> a row pair int16[ref_len] (intensity, dwell; -32768 = NaN): ur segments with
//                  the kit shift, uc / ul reversed, negative lengths dropped, zero lengths resolved,
//                  invalid-kmer ratio                        (uncalled4.py:118-201, common.py:298-329)
//   read_before    the read order of the comparison: by name, equal names in file / record order
//                                                            (run.py:698-700; readers.Uncalled4Reader.read)
//   position_count / position_fill   rows of the kept reads, in that order -> entries of a position
//                                                            (pack.cu: ncb_pack_rows_count / _fill_i16)
//
// Anything this path does not reproduce to the bit -- values that are not int16 codes, overlapping or
// very many ur segments -- or any malformed record sets a flag on the transcript; the caller then reads
// that transcript with the host reader (which also owns the error messages).
#pragma once
#include <stdint.h>

#ifndef NCB_HD
#ifdef __CUDACC__
#define NCB_HD __host__ __device__ __forceinline__
#else
#define NCB_HD inline
#endif
#endif

namespace ncb_decode {

enum { MAX_SEGMENTS = 48 };
enum {
    TX_OK = 0,
    TX_ERROR = 1,        // a member does not inflate, a record or its tags are malformed
    TX_NEEDS_HOST = 2    // well-formed, but outside what this path reproduces exactly
};
enum { SLOT_UNUSED = 0, SLOT_PRIMARY = 1, SLOT_KEPT = 2 };

struct BlockDesc {        // one BGZF member staged for the device
    uint64_t word_off;    // payload: first 32-bit word in the staging buffer (4-byte aligned, zero padded)
    uint32_t n_words;
    uint32_t cbytes;      // bytes of the raw DEFLATE payload
    uint32_t usize;       // ISIZE
    int32_t tx;
    int64_t dst_off;      // where its text goes
};

struct SpanDesc {         // one run of consecutive records of one transcript in one file
    int64_t text_off;     // text of the run's first member
    int64_t text_len;     // bytes of text of all its members
    int32_t in_block;     // first record, relative to text_off
    int32_t n_records;
    int32_t tx;
    int32_t file;
    int32_t first_slot;   // read slots [first_slot, first_slot + n_records)
    int32_t pad;
};

struct TxDesc {
    int64_t row_base;     // first int16 of the transcript's rows (slot s of the transcript: + s * ref_len)
    int64_t pos_base;     // first position of the transcript in the batch
    int32_t ref_len;
    int32_t first_slot;
    int32_t n_slots;
    int32_t pad;
};

struct ReadRec {
    int64_t name_off;                 // into the text
    int64_t ur_off, uc_off, ul_off;   // first element of the arrays
    uint32_t ur_n, uc_n, ul_n;
    uint8_t ur_sub, uc_sub, ul_sub, name_len;
    int32_t tx, file;
    int32_t state;                    // SLOT_*
    int32_t pad;
    uint64_t key0, key1;              // first 16 bytes of the name, big endian (shorter names zero padded)
};

NCB_HD uint32_t rd16u(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
NCB_HD uint32_t rd32u(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

NCB_HD int sub_size(uint8_t sub) {
    switch (sub) {
        case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        default: return 0;
    }
}

// element i of an integer aux array
NCB_HD int64_t aux_int(const uint8_t *data, uint8_t sub, uint32_t i) {
    switch (sub) {
        case 'c': return (int8_t)data[i];
        case 'C': return data[i];
        case 's': return (int16_t)rd16u(data + 2 * (size_t)i);
        case 'S': return (int64_t)rd16u(data + 2 * (size_t)i);
        case 'i': return (int32_t)rd32u(data + 4 * (size_t)i);
        default: return (int64_t)rd32u(data + 4 * (size_t)i);
    }
}

// ---- records of a span --------------------------------------------------------------------------
// The aux fields of a record: p .. end.  Fills the array fields of r; returns the bit mask of the signal
// tags found (7 = all), or -1 for malformed fields.
NCB_HD int find_signal_tags(const uint8_t *text, int64_t p, int64_t end, ReadRec &r) {
    int found = 0;
    while (p + 3 <= end) {
        const uint8_t t0 = text[p], t1 = text[p + 1], typ = text[p + 2];
        p += 3;
        switch (typ) {
            case 'A': case 'c': case 'C': p += 1; break;
            case 's': case 'S': p += 2; break;
            case 'i': case 'I': case 'f': p += 4; break;
            case 'Z': case 'H': {
                while (p < end && text[p] != 0) ++p;
                if (p >= end) return -1;
                ++p;
                break;
            }
            case 'B': {
                if (p + 5 > end) return -1;
                const uint8_t sub = text[p];
                const uint32_t n = rd32u(text + p + 1);
                const int size = sub_size(sub);
                const bool ours = t0 == 'u' && (t1 == 'r' || t1 == 'c' || t1 == 'l');
                if (size == 0) return -1;                   // the host reader accepts the same element types
                if ((uint64_t)n * (uint64_t)size > (uint64_t)(end - (p + 5))) return -1;
                if (ours) {
                    if (sub == 'f') return -1;              // the signal tags are integer arrays
                    if (t1 == 'r') { r.ur_off = p + 5; r.ur_n = n; r.ur_sub = sub; found |= 1; }
                    else if (t1 == 'c') { r.uc_off = p + 5; r.uc_n = n; r.uc_sub = sub; found |= 2; }
                    else { r.ul_off = p + 5; r.ul_n = n; r.ul_sub = sub; found |= 4; }
                }
                p += 5 + (int64_t)n * size;
                break;
            }
            default: return -1;
        }
    }
    return found;
}

// Walks the records of a span.  Every record takes a slot; secondary / supplementary alignments leave
// theirs SLOT_UNUSED (uncalled4.py:89).  Returns TX_OK, TX_ERROR or TX_NEEDS_HOST.
NCB_HD int walk_span(const uint8_t *text, const SpanDesc &sp, ReadRec *recs) {
    int64_t p = sp.text_off + sp.in_block;
    const int64_t end = sp.text_off + sp.text_len;
    for (int k = 0; k < sp.n_records; ++k) {
        ReadRec &r = recs[sp.first_slot + k];
        r.state = SLOT_UNUSED;
        r.tx = sp.tx;
        r.file = sp.file;
        if (p + 4 > end) return TX_ERROR;
        const int64_t block_size = (int32_t)rd32u(text + p);
        if (block_size < 32 || p + 4 + block_size > end) return TX_ERROR;
        const uint8_t *rec = text + p + 4;
        const int64_t rec_end = p + 4 + block_size;
        p = rec_end;
        const int l_read_name = rec[8];
        const int n_cigar = (int)rd16u(rec + 12);
        const int flag = (int)rd16u(rec + 14);
        const int64_t l_seq = (int32_t)rd32u(rec + 16);
        if (flag & (256 | 2048)) continue;
        if (l_seq < 0) return TX_ERROR;
        const int64_t fixed = (int64_t)l_read_name + 4 * (int64_t)n_cigar + (l_seq + 1) / 2 + l_seq;
        const int64_t body = rec_end - block_size + 32;        // first byte after the fixed-size fields
        if (fixed > rec_end - body) return TX_ERROR;
        r.name_off = body;
        // the host reader keeps names whole (up to 254 characters): so does the comparison here
        int name_len = l_read_name > 0 ? l_read_name - 1 : 0;
        // a name ends at its first NUL, like the C string the host reader copies
        for (int i = 0; i < name_len; ++i) {
            if (text[body + i] == 0) { name_len = i; break; }
            // the host path orders the names as decoded text: only for ASCII is that the byte order used here
            if (text[body + i] >= 0x80) return TX_NEEDS_HOST;
        }
        r.name_len = (uint8_t)name_len;
        uint64_t k0 = 0, k1 = 0;
        for (int i = 0; i < 8; ++i) k0 = (k0 << 8) | (i < name_len ? text[body + i] : 0);
        for (int i = 8; i < 16; ++i) k1 = (k1 << 8) | (i < name_len ? text[body + i] : 0);
        r.key0 = k0;
        r.key1 = k1;
        r.ur_n = r.uc_n = r.ul_n = 0;
        const int found = find_signal_tags(text, body + fixed, rec_end, r);
        if (found != 7) return TX_ERROR;
        r.state = SLOT_PRIMARY;
    }
    return TX_OK;
}

// ---- one read -> its rows --------------------------------------------------------------------------
struct Segments {            // per warp, shared memory on the device
    int32_t ref_start[MAX_SEGMENTS];
    int32_t sig_start[MAX_SEGMENTS + 1];   // first signal index of the segment; [n] = total
    int32_t n;
};

// reference position of signal index i
NCB_HD int64_t ref_of(const Segments &sg, int i) {
    int s = 0;
    while (s + 1 < sg.n && i >= sg.sig_start[s + 1]) ++s;
    return (int64_t)sg.ref_start[s] + (i - sg.sig_start[s]);
}

// W: the warp policy of inflate_core.cuh plus reduce_min / reduce_max / reduce_add over the lanes.
// Returns TX_OK / TX_ERROR / TX_NEEDS_HOST (the same value on every lane); on TX_OK r.state says whether the
// read is kept (invalid-kmer ratio below max_invalid_ratio, common.py:298-329 and run.py:711-718).
template <class W>
NCB_HD int decode_read(W &w, const uint8_t *text, ReadRec &r, int64_t ref_len, int kit_len, int kit_center,
                       double max_invalid_ratio, int16_t *irow, int16_t *drow, Segments &sg) {
    const int16_t NaN16 = (int16_t)-32768;
    for (int64_t i = w.lane(); i < ref_len; i += w.stride()) { irow[i] = NaN16; drow[i] = NaN16; }
    int status = TX_OK;
    if (w.leader()) {
        const int shift = kit_len - kit_center;
        const uint8_t *ur = text + r.ur_off, *ul = text + r.ul_off;
        if (r.ur_n < 2 || (r.ur_n & 1)) status = TX_ERROR;
        else if (r.ur_n / 2 > MAX_SEGMENTS) status = TX_NEEDS_HOST;
        else {
            // uncalled4.py:154-181
            const int64_t first = aux_int(ur, r.ur_sub, 0), last = aux_int(ur, r.ur_sub, r.ur_n - 1);
            int64_t signal_start = 0, prev_end = 0;
            int n = 0;
            for (uint32_t s = 0; s < r.ur_n && status == TX_OK; s += 2) {
                int64_t ref_start = aux_int(ur, r.ur_sub, s), ref_end = aux_int(ur, r.ur_sub, s + 1);
                if (ref_start == first) ref_start += shift;
                if (ref_end == last) ref_end -= kit_center - 1;
                ref_start -= shift;
                ref_end -= shift;
                const int64_t size = ref_end - ref_start;
                if (size < 0 || ref_start < 0 || ref_end > ref_len || signal_start + size > (int64_t)r.uc_n)
                    status = TX_ERROR;
                else if (n > 0 && ref_start < prev_end) status = TX_NEEDS_HOST;   // overlapping segments: order matters
                else {
                    sg.ref_start[n] = (int32_t)ref_start;
                    sg.sig_start[n] = (int32_t)signal_start;
                    ++n;
                    signal_start += size;
                    prev_end = ref_end;
                }
            }
            sg.n = n;
            sg.sig_start[n] = (int32_t)signal_start;
            // lengths: listed 3' -> 5'; negative = padding; zero = the dwell of the kmer before (uncalled4.py:139-147)
            if (status == TX_OK) {
                const int total = (int)signal_start;
                int k = 0, seg = 0;
                int64_t carry = 0;
                for (uint32_t i = r.ul_n; i-- > 0 && k < total;) {
                    int64_t v = aux_int(ul, r.ul_sub, i);
                    if (v < 0) continue;
                    if (v == 0 && k > 0) v = carry;
                    carry = v;
                    if (v > 32767) { status = TX_NEEDS_HOST; break; }
                    while (seg + 1 < n && k >= sg.sig_start[seg + 1]) ++seg;
                    drow[(int64_t)sg.ref_start[seg] + (k - sg.sig_start[seg])] = (int16_t)v;
                    ++k;
                }
                if (status == TX_OK && k < total) status = TX_ERROR;      // fewer lengths than the segments need
            }
        }
    }
    status = w.bcast(status);
    if (status != TX_OK) return status;
    w.sync();
    // currents: uc reversed, -32768 = NaN (uncalled4.py:103-107)
    {
        const uint8_t *uc = text + r.uc_off;
        const int total = sg.sig_start[sg.n];
        int bad = 0;
        for (int i = w.lane(); i < total; i += w.stride()) {
            const int64_t v = aux_int(uc, r.uc_sub, r.uc_n - 1 - (uint32_t)i);
            if (v < -32768 || v > 32767) { bad = 1; continue; }
            irow[ref_of(sg, i)] = (int16_t)v;
        }
        if (w.any(bad)) return TX_NEEDS_HOST;
    }
    w.sync();
    // common.get_reads_invalid_ratio on the finished row
    int64_t first = ref_len, last = -1;
    int valid = 0;
    for (int64_t i = w.lane(); i < ref_len; i += w.stride()) {
        if (irow[i] != NaN16) {
            if (i < first) first = i;
            if (i > last) last = i;
            ++valid;
        }
    }
    first = w.reduce_min(first);
    last = w.reduce_max(last);
    valid = w.reduce_add(valid);
    bool keep = false;
    if (valid > 0) {
        const double length = (double)(last - first + 1);
        keep = (length - (double)valid) / length < max_invalid_ratio;
    }
    if (w.leader()) r.state = keep ? SLOT_KEPT : SLOT_PRIMARY;
    return TX_OK;
}

// ---- read order ------------------------------------------------------------------------------------
// true when read a (slot ia) comes before read b (slot ib): by name, equal names by slot
NCB_HD bool read_before(const uint8_t *text, const ReadRec &a, int ia, const ReadRec &b, int ib) {
    if (a.key0 != b.key0) return a.key0 < b.key0;
    if (a.key1 != b.key1) return a.key1 < b.key1;
    const int n = a.name_len < b.name_len ? a.name_len : b.name_len;
    for (int i = 16; i < n; ++i) {
        const uint8_t x = text[a.name_off + i], y = text[b.name_off + i];
        if (x != y) return x < y;
    }
    if (a.name_len != b.name_len) return a.name_len < b.name_len;
    return ia < ib;
}

// ---- rows -> entries ---------------------------------------------------------------------------------
// order[first_slot + k] = slot of the k-th kept read of the transcript.
NCB_HD bool entry_of(const int16_t *rows_i, const int16_t *rows_d, const TxDesc &tx, int slot_in_tx, int p,
                     int motor_offset, int16_t &x, int16_t &y, int16_t &m) {
    const int64_t row = tx.row_base + (int64_t)slot_in_tx * tx.ref_len;
    x = rows_i[row + p];
    y = rows_d[row + p];
    m = (int16_t)-32768;
    if (motor_offset > 0 && p + motor_offset < tx.ref_len) m = rows_d[row + p + motor_offset];
    return x != (int16_t)-32768 || y != (int16_t)-32768 || m != (int16_t)-32768;
}

NCB_HD int position_count(const int16_t *rows_i, const int16_t *rows_d, const TxDesc &tx, const int32_t *order,
                          int n_kept, int p, int motor_offset) {
    int c = 0;
    for (int k = 0; k < n_kept; ++k) {
        int16_t x, y, m;
        c += entry_of(rows_i, rows_d, tx, order[tx.first_slot + k] - tx.first_slot, p, motor_offset, x, y, m) ? 1 : 0;
    }
    return c;
}

NCB_HD void position_fill(const int16_t *rows_i, const int16_t *rows_d, const TxDesc &tx, const int32_t *order,
                          int n_kept, int p, int motor_offset, const ReadRec *recs, const uint8_t *file_label,
                          int64_t at, int16_t *signal, uint8_t *label) {
    const int V = motor_offset > 0 ? 3 : 2;
    for (int k = 0; k < n_kept; ++k) {
        const int slot = order[tx.first_slot + k];
        int16_t x, y, m;
        if (entry_of(rows_i, rows_d, tx, slot - tx.first_slot, p, motor_offset, x, y, m)) {
            signal[V * at] = x;
            signal[V * at + 1] = y;
            if (V == 3) signal[V * at + 2] = m;
            label[at] = file_label[recs[slot].file];
            ++at;
        }
    }
}

}  // namespace ncb_decode
