// Native transcript readers (host code, no CUDA calls): SURVEY 8f row N1, north_star item (1).
//
//   ncb_bam_*               BGZF/BAM reader + Uncalled4 tag decoder.  Replaces, for one transcript,
//                           pysam.AlignmentFile.fetch/count + Uncalled4.get_data /
//                           _copy_signal_to_tensor / _resolve_skips
//                           (nanocompore/uncalled4.py:57-201) without the dense
//                           (positions, reads, 2) tensor: every read becomes one planar row pair
//                           intensity[ref_len], dwell[ref_len] (NaN = not covered), the layout
//                           the SQLite reader of the reference already stores per read
//                           (signal_data blobs, database.py:69-76; decoded at run.py:795-801).
//   ncb_rows_invalid_ratio  common.get_reads_invalid_ratio (common.py:298-329) on rows.
//
// The rows of all samples go to ncb_pack_rows_count / ncb_pack_rows_fill (pack.cu), which write
// the pinned CSR batch that ncb_compare streams to HBM.
//
// File access: the BAM is indexed once at open time (one sequential pass over the BGZF blocks,
// inflated `threads` at a time): for every reference the list of runs of consecutive records,
// each run addressed by a virtual offset (block file offset, offset inside the inflated block).
// A coordinate-sorted BAM has one run per reference, so reading a transcript inflates only the
// blocks that hold its records; no .bai is needed and unsorted files work too.
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <exception>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "ncb.h"

#include "bam_internal.h"

namespace {

// Sequential view of the inflated byte stream starting at a block, fed window by window.
struct Stream {
    ncb_bam *bam;
    int64_t next_off;                 // file offset of the next block to load
    static const int WINDOW = 128;    // most blocks inflated per refill (<= 8 MB of text)
    int window = 2;                   // grows geometrically: a short run inflates little past its end
    std::vector<Block> blocks;
    std::vector<std::vector<uint8_t>> data;
    size_t cur = 0;                   // current block of the window
    size_t pos = 0;                   // offset inside it
    bool failed = false;

    Stream(ncb_bam *b, int64_t off, int32_t in_block) : bam(b), next_off(off) {
        refill();
        pos = (size_t)in_block;
    }
    bool refill() {
        blocks.clear();
        data.clear();
        cur = 0;
        pos = 0;
        Block bl;
        while ((int)blocks.size() < window && next_off < bam->file_size) {
            if (!bam->block_at(next_off, bl)) { failed = true; return false; }
            blocks.push_back(bl);
            next_off += bl.csize;
        }
        if (blocks.empty()) return false;
        if (!bam->inflate_blocks(blocks, data)) { failed = true; return false; }
        window = window * 2 < WINDOW ? window * 2 : WINDOW;
        return true;
    }
    // virtual offset of the next byte: skips exhausted blocks first
    bool settle() {
        for (;;) {
            if (cur < data.size() && pos < data[cur].size()) return true;
            if (cur + 1 < data.size()) { ++cur; pos = 0; continue; }
            if (!refill()) return false;
            if (!data.empty() && !data[0].empty()) return true;
        }
    }
    bool at(int64_t &block_off, int32_t &in_block) {
        if (!settle()) return false;
        block_off = blocks[cur].file_off;
        in_block = (int32_t)pos;
        return true;
    }
    bool read(uint8_t *dst, size_t n) {
        while (n) {
            if (!settle()) return false;
            const size_t take = std::min(n, data[cur].size() - pos);
            memcpy(dst, data[cur].data() + pos, take);
            dst += take;
            pos += take;
            n -= take;
        }
        return true;
    }
};

bool build_index(ncb_bam *b) {
    Stream s(b, 0, 0);
    if (s.failed) return false;
    uint8_t w[8];
    if (!s.read(w, 8) || memcmp(w, "BAM\1", 4) != 0) {
        if (b->error.empty()) b->error = "not a BAM file (magic)";
        return false;
    }
    // every length field is checked before it sizes a buffer: a corrupt file must end in an error
    // message, not in an allocation failure
    const int32_t l_text = rdi32(w + 4);
    if (l_text < 0) { b->error = "corrupt BAM header (l_text)"; return false; }
    std::vector<uint8_t> buf;
    for (int64_t left = l_text; left > 0;) {          // the header text is skipped piecewise
        buf.resize((size_t)std::min<int64_t>(left, 1 << 20));
        if (!s.read(buf.data(), buf.size())) { b->error = "truncated BAM header"; return false; }
        left -= (int64_t)buf.size();
    }
    if (!s.read(w, 4)) { b->error = "truncated BAM header"; return false; }
    const int32_t n_ref = rdi32(w);
    if (n_ref < 0) { b->error = "corrupt BAM header (n_ref)"; return false; }
    for (int32_t i = 0; i < n_ref; ++i) {
        if (!s.read(w, 4)) { b->error = "truncated BAM reference list"; return false; }
        const int32_t l_name = rdi32(w);
        if (l_name < 1 || l_name > (1 << 20)) { b->error = "corrupt BAM reference list (l_name)"; return false; }
        buf.resize((size_t)l_name + 4);
        if (!s.read(buf.data(), buf.size())) { b->error = "truncated BAM reference list"; return false; }
        std::string name((const char *)buf.data(), (size_t)l_name - 1);
        b->ref_index.emplace(name, i);
        b->ref_names.push_back(name);
        b->ref_len.push_back(rdi32(buf.data() + l_name));
    }
    b->runs.assign((size_t)n_ref, {});
    b->ref_records.assign((size_t)n_ref, 0);
    int32_t prev = -2;
    for (;;) {
        int64_t boff;
        int32_t inb;
        if (!s.at(boff, inb)) break;   // clean end of file (or a failure flagged below)
        if (!s.read(w, 8)) { b->error = "truncated BAM record"; return false; }
        const int32_t block_size = rdi32(w), ref = rdi32(w + 4);
        if (block_size < 32 || block_size > NCB_BAM_MAX_RECORD) { b->error = "corrupt BAM record (block_size)"; return false; }
        buf.resize((size_t)block_size - 4);
        if (!s.read(buf.data(), buf.size())) { b->error = "truncated BAM record"; return false; }
        if (ref < 0 || ref >= n_ref) {
            ++b->n_unplaced;
            prev = -1;
            continue;
        }
        if (ref != prev) b->runs[ref].push_back(Run{boff, inb, 0, boff});
        ++b->runs[ref].back().n_records;
        // the read that fetched the record's last byte left the stream in the block that holds it
        b->runs[ref].back().end_block_off = s.blocks[s.cur].file_off;
        ++b->ref_records[ref];
        prev = ref;
    }
    return !s.failed;
}

// integer aux array -> int64 values
bool aux_ints(const uint8_t *p, const uint8_t *end, std::vector<int64_t> &out, const uint8_t **next) {
    if (p + 5 > end) return false;
    const char sub = (char)p[0];
    const uint32_t n = rd32(p + 1);
    p += 5;
    int size;
    switch (sub) {
        case 'c': case 'C': size = 1; break;
        case 's': case 'S': size = 2; break;
        case 'i': case 'I': case 'f': size = 4; break;
        default: return false;
    }
    if ((int64_t)(end - p) < (int64_t)n * size) return false;
    if (next) *next = p + (size_t)n * size;
    if (sub == 'f') return false;   // the signal tags are integer arrays
    out.resize(n);
    for (uint32_t i = 0; i < n; ++i, p += size) {
        switch (sub) {
            case 'c': out[i] = (int8_t)p[0]; break;
            case 'C': out[i] = p[0]; break;
            case 's': out[i] = (int16_t)rd16(p); break;
            case 'S': out[i] = rd16(p); break;
            case 'i': out[i] = rdi32(p); break;
            default: out[i] = rd32(p); break;
        }
    }
    return true;
}

// Walks the aux fields of a record and extracts the ur / uc / ul arrays.
bool find_signal_tags(const uint8_t *p, const uint8_t *end, std::vector<int64_t> &ur,
                      std::vector<int64_t> &uc, std::vector<int64_t> &ul, int &found) {
    found = 0;
    while (p + 3 <= end) {
        const char t0 = (char)p[0], t1 = (char)p[1], typ = (char)p[2];
        p += 3;
        switch (typ) {
            case 'A': case 'c': case 'C': p += 1; break;
            case 's': case 'S': p += 2; break;
            case 'i': case 'I': case 'f': p += 4; break;
            case 'Z': case 'H': {
                const void *z = memchr(p, 0, (size_t)(end - p));
                if (!z) return false;
                p = (const uint8_t *)z + 1;
                break;
            }
            case 'B': {
                std::vector<int64_t> *dst = nullptr;
                if (t0 == 'u' && t1 == 'r') dst = &ur;
                else if (t0 == 'u' && t1 == 'c') dst = &uc;
                else if (t0 == 'u' && t1 == 'l') dst = &ul;
                const uint8_t *next = nullptr;
                if (dst) {
                    if (!aux_ints(p, end, *dst, &next)) return false;
                    found |= dst == &ur ? 1 : (dst == &uc ? 2 : 4);
                } else {
                    if (p + 5 > end) return false;
                    const char sub = (char)p[0];
                    const uint32_t n = rd32(p + 1);
                    const int size = (sub == 'c' || sub == 'C') ? 1 : ((sub == 's' || sub == 'S') ? 2 : 4);
                    if ((uint64_t)n * (uint64_t)size > (uint64_t)(end - (p + 5))) return false;
                    next = p + 5 + (size_t)n * size;
                }
                p = next;
                break;
            }
            default: return false;
        }
    }
    return true;
}

}  // namespace

extern "C" {

int ncb_bam_open(ncb_bam **out, const char *path, int threads) {
    if (!out || !path) return NCB_ERR_INVALID;
    *out = nullptr;
    ncb_bam *b = new (std::nothrow) ncb_bam();
    if (!b) return NCB_ERR_NOMEM;
    b->path = path;
    b->threads = threads < 1 ? 1 : threads;
    const int fd = open(path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {
        if (fd >= 0) close(fd);
        delete b;
        return NCB_ERR_INVALID;
    }
    b->file_size = (int64_t)st.st_size;
    if (b->file_size > 0) {
        void *m = mmap(nullptr, (size_t)b->file_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) {
            close(fd);
            delete b;
            return NCB_ERR_NOMEM;
        }
        b->map = (const uint8_t *)m;
    }
    close(fd);
    *out = b;   // returned even on a parse failure so that the caller can read the message
    // nothing is thrown across the C ABI
    try {
        if (!build_index(b)) {
            if (b->error.empty()) b->error = "cannot parse BAM";
            return NCB_ERR_INVALID;
        }
    } catch (const std::bad_alloc &) {
        b->error = "out of memory while indexing the BAM";
        return NCB_ERR_NOMEM;
    } catch (const std::exception &e) {
        b->error = std::string("cannot parse BAM: ") + e.what();
        return NCB_ERR_INVALID;
    }
    return NCB_OK;
}

void ncb_bam_close(ncb_bam *b) {
    if (!b) return;
    if (b->map) munmap(const_cast<uint8_t *>(b->map), (size_t)b->file_size);
    delete b;
}

const char *ncb_bam_last_error(const ncb_bam *b) { return b ? b->error.c_str() : "null handle"; }

int32_t ncb_bam_n_references(const ncb_bam *b) { return b ? (int32_t)b->ref_names.size() : 0; }

const char *ncb_bam_reference_name(const ncb_bam *b, int32_t i) {
    return (b && i >= 0 && i < (int32_t)b->ref_names.size()) ? b->ref_names[i].c_str() : nullptr;
}

int64_t ncb_bam_reference_length(const ncb_bam *b, int32_t i) {
    return (b && i >= 0 && i < (int32_t)b->ref_len.size()) ? b->ref_len[i] : -1;
}

int64_t ncb_bam_reference_records(const ncb_bam *b, int32_t i) {
    return (b && i >= 0 && i < (int32_t)b->ref_records.size()) ? b->ref_records[i] : -1;
}

int32_t ncb_bam_reference_index(const ncb_bam *b, const char *name) {
    if (!b || !name) return -1;
    auto it = b->ref_index.find(name);
    return it == b->ref_index.end() ? -1 : it->second;
}

int64_t ncb_bam_read_uncalled4(ncb_bam *b, int32_t ref, int64_t ref_len, int32_t kit_len,
                               int32_t kit_center, float *intensity, float *dwell, char *names,
                               int32_t name_stride, int64_t max_reads) {
    if (!b || ref < 0 || ref >= (int32_t)b->runs.size() || ref_len < 0 || !names || name_stride < 2 ||
        (max_reads > 0 && ref_len > 0 && (!intensity || !dwell)))
        return NCB_ERR_INVALID;
    b->error.clear();
    try {
    const float NaN = nanf("");
    const int shift = kit_len - kit_center;
    int64_t n_out = 0;
    std::vector<uint8_t> rec;
    std::vector<int64_t> ur, uc, ul;
    std::vector<double> lens;
    uint8_t w[4];
    for (const Run &run : b->runs[ref]) {
        Stream s(b, run.block_off, run.in_block);
        if (s.failed) return NCB_ERR_INVALID;
        for (int64_t k = 0; k < run.n_records; ++k) {
            if (!s.read(w, 4)) { b->error = "truncated BAM record"; return NCB_ERR_INVALID; }
            const int32_t block_size = rdi32(w);
            if (block_size < 32 || block_size > NCB_BAM_MAX_RECORD) {
                b->error = "corrupt BAM record (block_size)";
                return NCB_ERR_INVALID;
            }
            rec.resize((size_t)block_size);
            if (!s.read(rec.data(), rec.size())) {
                b->error = "truncated BAM record";
                return NCB_ERR_INVALID;
            }
            const uint8_t *r = rec.data();
            const int l_read_name = r[8];
            const int n_cigar = rd16(r + 12);
            const int flag = rd16(r + 14);
            const int32_t l_seq = rdi32(r + 16);
            if (flag & (256 | 2048)) continue;   // secondary / supplementary (uncalled4.py:89)
            if (n_out >= max_reads) { b->error = "more primary reads than the caller's capacity"; return NCB_ERR_INVALID; }
            const uint8_t *p = r + 32;
            const uint8_t *end = r + block_size;
            if (l_seq < 0) { b->error = "corrupt BAM record (l_seq)"; return NCB_ERR_INVALID; }
            const size_t fixed = (size_t)l_read_name + 4u * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
            if (fixed > (size_t)(end - p)) { b->error = "corrupt BAM record"; return NCB_ERR_INVALID; }
            char *name = names + n_out * name_stride;
            memset(name, 0, (size_t)name_stride);
            memcpy(name, p, (size_t)std::min(l_read_name > 0 ? l_read_name - 1 : 0, name_stride - 1));
            int found = 0;
            if (!find_signal_tags(p + fixed, end, ur, uc, ul, found)) {
                b->error = "corrupt aux fields in read " + std::string(name);
                return NCB_ERR_INVALID;
            }
            if (found != 7) {   // pysam's get_tag raises KeyError
                b->error = "read " + std::string(name) + " lacks the Uncalled4 tags ur/uc/ul";
                return NCB_ERR_INVALID;
            }
            float *irow = intensity + n_out * ref_len;
            float *drow = dwell + n_out * ref_len;
            for (int64_t i = 0; i < ref_len; ++i) { irow[i] = NaN; drow[i] = NaN; }
            // uncalled4.py:139-147: both arrays are listed 3' -> 5'; negative lengths are
            // padding; a zero length repeats the previous kmer's dwell (_resolve_skips)
            std::reverse(uc.begin(), uc.end());
            lens.clear();
            for (size_t i = ul.size(); i-- > 0;)
                if (ul[i] >= 0) lens.push_back((double)ul[i]);
            for (size_t i = 1; i < lens.size(); ++i)
                if (lens[i] == 0.0) lens[i] = lens[i - 1];
            if (ur.size() < 2 || (ur.size() & 1)) {
                b->error = "read " + std::string(name) + ": odd ur tag";
                return NCB_ERR_INVALID;
            }
            int64_t signal_start = 0;
            for (size_t sgm = 0; sgm < ur.size(); sgm += 2) {
                int64_t ref_start = ur[sgm], ref_end = ur[sgm + 1];
                // uncalled4.py:154-181
                if (ref_start == ur.front()) ref_start += shift;
                if (ref_end == ur.back()) ref_end -= kit_center - 1;
                ref_start -= shift;
                ref_end -= shift;
                const int64_t size = ref_end - ref_start;
                if (size < 0 || ref_start < 0 || ref_end > ref_len ||
                    signal_start + size > (int64_t)uc.size() || signal_start + size > (int64_t)lens.size()) {
                    b->error = "read " + std::string(name) + ": signal does not fit its ur segments";
                    return NCB_ERR_INVALID;
                }
                for (int64_t i = 0; i < size; ++i) {
                    const int64_t cur = uc[signal_start + i];
                    // -32768 is Uncalled4's NaN (uncalled4.py:103-107)
                    irow[ref_start + i] = cur == -32768 ? NaN : (float)cur;
                    const float len = (float)lens[signal_start + i];
                    drow[ref_start + i] = len == -32768.0f ? NaN : len;
                }
                signal_start += size;
            }
            ++n_out;
        }
    }
    return n_out;
    } catch (const std::bad_alloc &) {       // nothing is thrown across the C ABI
        b->error = "out of memory while reading the BAM";
        return NCB_ERR_NOMEM;
    } catch (const std::exception &e) {
        b->error = std::string("cannot read the BAM: ") + e.what();
        return NCB_ERR_INVALID;
    }
}

int ncb_rows_invalid_ratio(const float *const *intensity, int64_t n_rows, int64_t n_positions,
                           double *ratios, int threads) {
    if ((n_rows > 0 && (!intensity || !ratios)) || n_rows < 0 || n_positions < 0) return NCB_ERR_INVALID;
    auto work = [=](int64_t lo, int64_t hi) {
        for (int64_t r = lo; r < hi; ++r) {
            const float *row = intensity[r];
            int64_t first = -1, last = -1, valid = 0;
            for (int64_t p = 0; p < n_positions; ++p)
                if (row[p] == row[p]) {
                    if (first < 0) first = p;
                    last = p;
                    ++valid;
                }
            if (valid == 0) {
                ratios[r] = nan("");   // numpy raises on the empty max(): such a read is never kept
            } else {
                const double length = (double)(last - first + 1);
                ratios[r] = (length - (double)valid) / length;
            }
        }
    };
    const int nt = (int)std::min<int64_t>(std::max(threads, 1), std::max<int64_t>(n_rows, 1));
    if (nt <= 1) {
        work(0, n_rows);
    } else {
        std::vector<std::thread> pool;
        const int64_t chunk = (n_rows + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) {
            const int64_t lo = t * chunk, hi = std::min(n_rows, lo + chunk);
            if (lo >= hi) continue;
            try {
                pool.emplace_back(work, lo, hi);
            } catch (const std::exception &) {
                work(lo, hi);
            }
        }
        for (auto &th : pool) th.join();
    }
    return NCB_OK;
}

}  // extern "C"
