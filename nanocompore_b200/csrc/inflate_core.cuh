// DEFLATE decoder for BGZF members (RFC 1951), written for one warp per member: lane 0 walks the bit
// stream and turns it into commands (literal / copy), the warp executes the commands (inflate_warp in
// decode.cu).  Everything that is sequential lives here as host/device functions on plain arrays, so
// the same code is compiled for the host by tests/csrc/inflate_host.cpp and checked against zlib on
// ordinary, fixed-Huffman, stored and damaged streams (tests/test_inflate_host.py).
//
// Replaces zlib's inflate as used by the reference's BAM access (pysam/htslib bgzf.c behind
// uncalled4.py:76-77) and by this library's own host reader (reader.cu: inflate_blocks).  The
// acceptance rules are zlib's (inftrees.c / inflate.c): over-subscribed or incomplete code sets,
// missing end-of-block code, repeats without a previous length, invalid symbols and distances
// beyond the start of the member are errors; the stream has to end exactly at ISIZE bytes.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NCB_HD __host__ __device__ __forceinline__
#else
#define NCB_HD inline
#endif

namespace ncb_inflate {

enum { LIT_BITS = 10, DIST_BITS = 8, MAX_CMDS = 32, MAX_LIT = 288, MAX_DIST = 32 };

enum {
    INF_OK = 0,
    INF_ERR_INPUT = 1,    // the stream runs past its compressed bytes
    INF_ERR_BTYPE = 2,    // block type 3
    INF_ERR_STORED = 3,   // stored block: LEN / NLEN do not match
    INF_ERR_HEADER = 4,   // dynamic block: bad counts, bad code-length code, bad repeat, no end-of-block code
    INF_ERR_CODE = 5,     // invalid literal/length or distance code
    INF_ERR_DIST = 6,     // distance beyond the start of the member
    INF_ERR_OUTPUT = 7    // more or fewer bytes than ISIZE
};

// per-warp tables (shared memory on the device)
struct Tables {
    uint16_t lit_lut[1 << LIT_BITS];    // sym | len << 9; 0: code longer than LIT_BITS or unused
    uint16_t dist_lut[1 << DIST_BITS];  // sym | len << 5 | 0x8000 (so that a valid entry is never 0)
    uint16_t lit_sorted[MAX_LIT];       // symbols ordered by (code length, symbol): the bit-by-bit path
    uint16_t lit_count[16];
    uint16_t dist_sorted[MAX_DIST];
    uint16_t dist_count[16];
    uint8_t lens[MAX_LIT + MAX_DIST];   // code lengths of the block: literal/length codes, then distance codes
    uint16_t code[MAX_LIT + MAX_DIST];  // their canonical codes, most significant bit first
    uint16_t cmd_len[MAX_CMDS];         // command queue: bytes the command produces (1 = literal)
    uint16_t cmd_val[MAX_CMDS];         // literal byte, or the distance of a copy - 1
    uint16_t cmd_off[MAX_CMDS + 1];     // first output byte of the command, relative to the batch; [n] = bytes of the batch
    int32_t n_lit, n_dist;              // codes of the current block
};

// ---- bit reader: 32-bit words of the member's payload, least significant bit first ---------------
struct BitReader {
    const uint32_t *words;   // 4-byte aligned; words past n_words read as 0
    uint32_t n_words;
    uint32_t next;           // next word to enter the buffer
    uint32_t ahead;          // words[next], loaded when the word before it entered the buffer: the load has a
                             // refill interval (~4 symbols) to arrive instead of sitting on the walk's chain
    uint64_t buf;
    int cnt;
    int64_t limit_bits;      // bits of the compressed payload

    NCB_HD void init(const uint32_t *w, uint32_t nw, int64_t bytes) {
        words = w; n_words = nw; next = 0; buf = 0; cnt = 0; limit_bits = bytes * 8;
        ahead = nw ? w[0] : 0u;
    }
    NCB_HD void need(int n) {        // n <= 32
        if (cnt < n) {
            buf |= (uint64_t)ahead << cnt;
            cnt += 32;
            ++next;
            ahead = next < n_words ? words[next] : 0u;
        }
    }
    NCB_HD uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    NCB_HD void drop(int n) { buf >>= n; cnt -= n; }
    NCB_HD uint32_t take(int n) {
        need(n);
        const uint32_t v = peek(n);
        drop(n);
        return v;
    }
    NCB_HD int64_t consumed() const { return (int64_t)next * 32 - cnt; }
    NCB_HD void align_to_byte() { drop(cnt & 7); }
    NCB_HD void seek_byte(int64_t byte) {      // continue at a byte position of the payload
        next = (uint32_t)(byte >> 2);
        ahead = next < n_words ? words[next] : 0u;
        buf = 0;
        cnt = 0;
        const int rest = (int)(byte & 3) * 8;
        if (rest) { need(rest); drop(rest); }
    }
    NCB_HD bool overrun() const { return consumed() > limit_bits; }
};

NCB_HD uint32_t bit_reverse(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// Canonical codes of lens[0, n): count[], sorted[], code[].  zlib's acceptance (inftrees.c): an
// over-subscribed set is an error; an incomplete one too unless it is a single code of length 1
// (`allow_single`: literal/length and distance codes, not the code-length code) or empty.
// Returns 0, or -1 for a set zlib rejects.
NCB_HD int build_canonical(const uint8_t *lens, int n, uint16_t *count, uint16_t *sorted, uint16_t *code,
                           bool allow_single) {
    for (int l = 0; l < 16; ++l) count[l] = 0;
    for (int s = 0; s < n; ++s) ++count[lens[s]];
    int max = 15;
    while (max >= 1 && count[max] == 0) --max;
    if (max == 0) {                 // no codes at all: every lookup fails (zlib builds such a table)
        count[0] = 0;
        return 0;
    }
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return -1;
    }
    if (left > 0 && (!allow_single || max != 1)) return -1;
    uint16_t offs[16], next_code[16];
    offs[1] = 0;
    next_code[1] = 0;
    uint32_t c = 0;
    for (int l = 1; l < 15; ++l) {
        offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        c = (c + count[l]) << 1;
        next_code[l + 1] = (uint16_t)c;
    }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l) {
            sorted[offs[l]++] = (uint16_t)s;
            code[s] = next_code[l]++;
        }
    }
    count[0] = 0;
    return 0;
}

// Lookup-table entries of the symbols lane, lane + stride, ... (the table was cleared before).
NCB_HD void fill_lit_lut(Tables &t, int lane, int stride) {
    for (int s = lane; s < t.n_lit; s += stride) {
        const int l = t.lens[s];
        if (l == 0 || l > LIT_BITS) continue;
        const uint32_t rev = bit_reverse(t.code[s], l);
        const uint16_t e = (uint16_t)(s | (l << 9));
        for (uint32_t k = rev; k < (1u << LIT_BITS); k += 1u << l) t.lit_lut[k] = e;
    }
}
NCB_HD void fill_dist_lut(Tables &t, int lane, int stride) {
    for (int s = lane; s < t.n_dist; s += stride) {
        const int l = t.lens[t.n_lit + s];
        if (l == 0 || l > DIST_BITS) continue;
        const uint32_t rev = bit_reverse(t.code[t.n_lit + s], l);
        const uint16_t e = (uint16_t)(s | (l << 5) | 0x8000);
        for (uint32_t k = rev; k < (1u << DIST_BITS); k += 1u << l) t.dist_lut[k] = e;
    }
}

// bit-by-bit canonical decode (codes longer than the table's index, and the error path); -1: invalid code
NCB_HD int decode_slow(BitReader &br, const uint16_t *count, const uint16_t *sorted) {
    br.need(15);
    uint32_t bits = br.peek(15);
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; ++l) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[l];
        if (code - c < first) {
            br.drop(l);
            return sorted[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

NCB_HD int decode_lit(BitReader &br, const Tables &t) {
    br.need(LIT_BITS);
    const uint16_t e = t.lit_lut[br.peek(LIT_BITS)];
    if (e) {
        br.drop(e >> 9);
        return e & 0x1ff;
    }
    return decode_slow(br, t.lit_count, t.lit_sorted);
}
NCB_HD int decode_dist(BitReader &br, const Tables &t) {
    br.need(DIST_BITS);
    const uint16_t e = t.dist_lut[br.peek(DIST_BITS)];
    if (e) {
        br.drop((e >> 5) & 0xf);
        return e & 0x1f;
    }
    return decode_slow(br, t.dist_count, t.dist_sorted);
}

// lengths of the fixed codes (RFC 1951 3.2.6)
NCB_HD void fixed_lengths(Tables &t) {
    for (int s = 0; s < 144; ++s) t.lens[s] = 8;
    for (int s = 144; s < 256; ++s) t.lens[s] = 9;
    for (int s = 256; s < 280; ++s) t.lens[s] = 7;
    for (int s = 280; s < 288; ++s) t.lens[s] = 8;
    for (int s = 0; s < 32; ++s) t.lens[288 + s] = 5;
    t.n_lit = 288;
    t.n_dist = 32;      // symbols 30 and 31 take part in the code and are invalid when decoded
}

// Header of a dynamic block: fills t.lens / n_lit / n_dist.  INF_OK or an error.  Uses lit_lut as
// scratch for the code-length code (the caller rebuilds the tables afterwards).
NCB_HD int read_dynamic_header(BitReader &br, Tables &t) {
    const int nlen = (int)br.take(5) + 257, ndist = (int)br.take(5) + 1, ncode = (int)br.take(4) + 4;
    if (nlen > 286 || ndist > 30) return INF_ERR_HEADER;
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19];
    for (int i = 0; i < 19; ++i) cl[i] = 0;
    for (int i = 0; i < ncode; ++i) cl[order[i]] = (uint8_t)br.take(3);
    uint16_t count[16], sorted[19], code[19];
    if (build_canonical(cl, 19, count, sorted, code, false) != 0) return INF_ERR_HEADER;
    int i = 0;
    while (i < nlen + ndist) {
        const int sym = decode_slow(br, count, sorted);
        if (sym < 0) return INF_ERR_HEADER;
        if (sym < 16) {
            t.lens[i++] = (uint8_t)sym;
        } else {
            int prev = 0, rep;
            if (sym == 16) {
                if (i == 0) return INF_ERR_HEADER;
                prev = t.lens[i - 1];
                rep = 3 + (int)br.take(2);
            } else if (sym == 17) {
                rep = 3 + (int)br.take(3);
            } else {
                rep = 11 + (int)br.take(7);
            }
            if (i + rep > nlen + ndist) return INF_ERR_HEADER;
            while (rep--) t.lens[i++] = (uint8_t)prev;
        }
        if (br.overrun()) return INF_ERR_INPUT;
    }
    if (t.lens[256] == 0) return INF_ERR_HEADER;     // no end-of-block code
    t.n_lit = nlen;
    t.n_dist = ndist;
    return INF_OK;
}

// canonical codes of the block's two alphabets (sequential part of the table set-up)
NCB_HD int build_block_codes(Tables &t) {
    if (build_canonical(t.lens, t.n_lit, t.lit_count, t.lit_sorted, t.code, true) != 0) return INF_ERR_HEADER;
    if (build_canonical(t.lens + t.n_lit, t.n_dist, t.dist_count, t.dist_sorted, t.code + t.n_lit, true) != 0)
        return INF_ERR_HEADER;
    return INF_OK;
}

// Decodes up to MAX_CMDS commands of a compressed block into the queue.  `produced` = bytes already
// written to the member's output, `usize` its ISIZE.  Returns the number of commands; *end_of_block is
// set when the end-of-block code was read, *err on an error.
NCB_HD int decode_commands(BitReader &br, Tables &t, int64_t produced, int64_t usize, bool *end_of_block, int *err) {
    int n = 0;
    const int64_t start = produced;
    *end_of_block = false;
    *err = INF_OK;
    t.cmd_off[0] = 0;
    while (n < MAX_CMDS) {
        const int sym = decode_lit(br, t);
        if (sym < 0) { *err = INF_ERR_CODE; break; }
        if (sym < 256) {
            if (produced + 1 > usize) { *err = INF_ERR_OUTPUT; break; }
            t.cmd_len[n] = 1;
            t.cmd_val[n] = (uint16_t)sym;
            produced += 1;
            t.cmd_off[++n] = (uint16_t)(produced - start);
        } else if (sym == 256) {
            *end_of_block = true;
            break;
        } else {
            if (sym > 285) { *err = INF_ERR_CODE; break; }
            // base values and extra bits of RFC 1951 3.2.5 in closed form (tables in local memory would be
            // initialised at every call): lengths 3..10 and 258 have no extra bits, from code 265 on four codes
            // share an extra-bit count; distances 1..4 have none, then two codes share one
            const int li = sym - 257;
            const int le = (li < 8 || li == 28) ? 0 : (li >> 2) - 1;
            const int lb = li == 28 ? 258 : (li < 8 ? li + 3 : 3 + ((4 + (li & 3)) << le));
            const int len = lb + (int)br.take(le);
            const int ds = decode_dist(br, t);
            if (ds < 0 || ds > 29) { *err = INF_ERR_CODE; break; }
            const int de = ds < 4 ? 0 : (ds >> 1) - 1;
            const int db = ds < 4 ? ds + 1 : 1 + ((2 + (ds & 1)) << de);
            const int dist = db + (int)br.take(de);
            if ((int64_t)dist > produced) { *err = INF_ERR_DIST; break; }
            if (produced + len > usize) { *err = INF_ERR_OUTPUT; break; }
            t.cmd_len[n] = (uint16_t)len;
            t.cmd_val[n] = (uint16_t)(dist - 1);      // a distance is at most 32768
            produced += len;
            t.cmd_off[++n] = (uint16_t)(produced - start);
        }
        if (br.overrun()) { *err = INF_ERR_INPUT; break; }
    }
    // the end-of-block code, too, has to lie inside the payload (zeros are read behind it, and seven
    // zero bits are the end-of-block code of the fixed code)
    if (*err == INF_OK && br.overrun()) *err = INF_ERR_INPUT;
    return n;
}


// decode_commands for a stretch of the stream where nothing unusual happens, with what the walk of one lane
// spends most of its instructions on taken out of the loop: the end of the payload is checked once per
// batch of commands instead of once per symbol (behind the payload the reader delivers zeros, and nothing is
// written to the output before the batch is accepted), offsets and bounds are 32-bit.  Returns the number of
// commands, or -1 when anything is off (invalid code, distance, output bound, payload overrun): the caller
// puts the reader back to the start of the batch and lets decode_commands find out what, so errors are
// reported exactly as before.
NCB_HD int decode_commands_fast(BitReader &br, Tables &t, int64_t produced, int64_t usize, bool *end_of_block) {
    const int64_t room64 = usize - produced;
    const int room = room64 > 0x7fff0000 ? 0x7fff0000 : (int)room64;      // bytes the output still takes
    const int have = produced > 0x10000 ? 0x10000 : (int)produced;        // bytes a distance may reach back, capped
    int n = 0, off = 0;
    *end_of_block = false;
    t.cmd_off[0] = 0;
    while (n < MAX_CMDS) {
        const int sym = decode_lit(br, t);
        if (sym < 256) {
            if (sym < 0 || off >= room) return -1;
            t.cmd_len[n] = 1;
            t.cmd_val[n] = (uint16_t)sym;
            t.cmd_off[++n] = (uint16_t)++off;
        } else if (sym == 256) {
            *end_of_block = true;
            break;
        } else {
            if (sym > 285) return -1;
            const int li = sym - 257;
            const int le = (li < 8 || li == 28) ? 0 : (li >> 2) - 1;
            const int lb = li == 28 ? 258 : (li < 8 ? li + 3 : 3 + ((4 + (li & 3)) << le));
            const int len = lb + (int)br.take(le);
            const int ds = decode_dist(br, t);
            if (ds < 0 || ds > 29) return -1;
            const int de = ds < 4 ? 0 : (ds >> 1) - 1;
            const int db = ds < 4 ? ds + 1 : 1 + ((2 + (ds & 1)) << de);
            const int dist = db + (int)br.take(de);
            if (dist > have + off || off + len > room) return -1;
            t.cmd_len[n] = (uint16_t)len;
            t.cmd_val[n] = (uint16_t)(dist - 1);
            off += len;
            t.cmd_off[++n] = (uint16_t)off;
        }
    }
    return br.overrun() ? -1 : n;
}

// ---- one member, by a "warp" W ------------------------------------------------------------------
// W provides lane() / stride() (the lanes share strided loops), leader() (the lane that walks the bit
// stream), bcast(int) (the leader's value on every lane), sync() (orders the lanes' accesses to the
// tables and to the output) and load_out(p) (a byte of the output written earlier by any lane).  On
// the device W is a CUDA warp (decode.cu); on the host one lane with stride 1 (the test harness).
// 32 commands of 258 bytes fit the 16-bit offsets of the queue.
template <class W>
NCB_HD int inflate_member(W &w, Tables &t, const uint32_t *words, uint32_t n_words, int64_t cbytes,
                          uint8_t *out, int64_t usize) {
    BitReader br;
    br.init(words, n_words, cbytes);
    const uint8_t *in_bytes = reinterpret_cast<const uint8_t *>(words);
    int64_t produced = 0;
    int err = INF_OK, final_block = 0;
    while (!final_block && !err) {
        int btype = 0;
        if (w.leader()) {
            final_block = (int)br.take(1);
            btype = (int)br.take(2);
            if (br.overrun()) err = INF_ERR_INPUT;
        }
        final_block = w.bcast(final_block);
        btype = w.bcast(btype);
        err = w.bcast(err);
        if (err) break;
        if (btype == 0) {
            int len = 0, from = 0;
            if (w.leader()) {
                br.align_to_byte();
                len = (int)br.take(16);
                const int nlen = (int)br.take(16);
                from = (int)(br.consumed() >> 3);
                if ((len ^ nlen) != 0xffff) err = INF_ERR_STORED;
                else if ((int64_t)from + len > cbytes) err = INF_ERR_INPUT;
                else if (produced + len > usize) err = INF_ERR_OUTPUT;
                else br.seek_byte((int64_t)from + len);
            }
            len = w.bcast(len);
            from = w.bcast(from);
            err = w.bcast(err);
            if (err) break;
            for (int i = w.lane(); i < len; i += w.stride()) out[produced + i] = in_bytes[from + i];
            produced += len;
            w.sync();
            continue;
        }
        if (btype == 3) { err = INF_ERR_BTYPE; break; }
        if (w.leader()) {
            if (btype == 1) fixed_lengths(t);
            else err = read_dynamic_header(br, t);
            if (!err) err = build_block_codes(t);
        }
        err = w.bcast(err);
        if (err) break;
        for (int i = w.lane(); i < (1 << LIT_BITS); i += w.stride()) t.lit_lut[i] = 0;
        for (int i = w.lane(); i < (1 << DIST_BITS); i += w.stride()) t.dist_lut[i] = 0;
        w.sync();
        fill_lit_lut(t, w.lane(), w.stride());
        fill_dist_lut(t, w.lane(), w.stride());
        w.sync();
        int eob = 0;
        while (!eob && !err) {
            int n = 0;
            if (w.leader()) {
                bool e = false;
                const BitReader at_start = br;
                n = decode_commands_fast(br, t, produced, usize, &e);
                if (n < 0) {
                    br = at_start;
                    n = decode_commands(br, t, produced, usize, &e, &err);
                }
                eob = e ? 1 : 0;
            }
            n = w.bcast(n);
            eob = w.bcast(eob);
            err = w.bcast(err);
            if (err) break;
            w.sync();
            // literals first: every lane its own bytes
            int copies = 0;
            for (int j = w.lane(); j < n; j += w.stride()) {
                if (t.cmd_len[j] == 1) out[produced + t.cmd_off[j]] = (uint8_t)t.cmd_val[j];
                else copies = 1;
            }
            if (w.any(copies)) {
                w.sync();
                // copies in stream order: byte i of a copy is the byte dist behind it, i.e. byte i mod dist of
                // the dist bytes before its start -- all written before this command
                for (int j = 0; j < n; ++j) {
                    const int len = t.cmd_len[j];
                    if (len == 1) continue;
                    const int dist = (int)t.cmd_val[j] + 1;
                    uint8_t *dst = out + produced + t.cmd_off[j];
                    const uint8_t *src = dst - dist;
                    for (int i = w.lane(); i < len; i += w.stride()) dst[i] = w.load_out(src + (dist >= len ? i : i % dist));
                    w.sync();
                }
            }
            produced += t.cmd_off[n];
            w.sync();
        }
    }
    if (!err && produced != usize) err = INF_ERR_OUTPUT;
    return err;
}

// the host's "warp": one lane
struct SerialWarp {
    NCB_HD int lane() const { return 0; }
    NCB_HD int stride() const { return 1; }
    NCB_HD bool leader() const { return true; }
    NCB_HD int bcast(int v) const { return v; }
    NCB_HD bool any(int v) const { return v != 0; }
    NCB_HD void sync() const {}
    NCB_HD uint8_t load_out(const uint8_t *p) const { return *p; }
    NCB_HD int64_t reduce_min(int64_t v) const { return v; }
    NCB_HD int64_t reduce_max(int64_t v) const { return v; }
    NCB_HD int reduce_add(int v) const { return v; }
};

}  // namespace ncb_inflate
