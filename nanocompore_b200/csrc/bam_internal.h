// Internals of the native BAM reader shared by reader.cu (host decoder) and decode.cu (staging of the
// device-side decoder): the handle with its index of runs, BGZF block access.
#pragma once
#include <errno.h>
#include <stdint.h>
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

// largest record accepted (a record holds one read: name, CIGAR, sequence, qualities and the signal tags)
#define NCB_BAM_MAX_RECORD (1 << 30)

struct Run {
    int64_t block_off;   // file offset of the BGZF block that holds the first record of the run
    int32_t in_block;    // offset of that record inside the inflated block
    int64_t n_records;
    int64_t end_block_off;   // file offset of the block that holds the last byte of the run's last record
};

struct Block {
    int64_t file_off;
    int32_t csize;       // whole BGZF member
    int32_t usize;       // ISIZE
};

static inline uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline int32_t rdi32(const uint8_t *p) { return (int32_t)rd32(p); }

// Message of the last failure on a handle, per calling thread.  Transcripts may be read from one
// handle by several threads at once (the tables below are read-only after ncb_bam_open); each call
// clears and sets the text of ITS thread, and ncb_bam_last_error returns the calling thread's: a
// failure on one transcript is not wiped by a call that starts on another meanwhile.
struct ErrorText {
    mutable std::mutex mu;
    std::unordered_map<std::thread::id, std::string> by_thread;   // node-based: the strings stay in place
    ErrorText &operator=(const std::string &s) {
        std::lock_guard<std::mutex> g(mu);
        by_thread[std::this_thread::get_id()] = s;
        return *this;
    }
    ErrorText &operator=(const char *s) { return *this = std::string(s); }
    void clear() {
        std::lock_guard<std::mutex> g(mu);
        auto it = by_thread.find(std::this_thread::get_id());
        if (it != by_thread.end()) it->second.clear();
    }
    bool empty() const {
        std::lock_guard<std::mutex> g(mu);
        auto it = by_thread.find(std::this_thread::get_id());
        return it == by_thread.end() || it->second.empty();
    }
    const char *c_str() const {
        std::lock_guard<std::mutex> g(mu);
        auto it = by_thread.find(std::this_thread::get_id());
        return it == by_thread.end() ? "" : it->second.c_str();
    }
};

struct ncb_bam {
    std::string path;
    const uint8_t *map = nullptr;   // the whole file, read-only mapping
    int64_t file_size = 0;
    int threads = 1;
    ErrorText error;
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_len;
    std::unordered_map<std::string, int32_t> ref_index;
    std::vector<std::vector<Run>> runs;      // per reference
    std::vector<int64_t> ref_records;        // per reference, every record (pysam count())
    int64_t n_unplaced = 0;                  // records with refID < 0

    // ---- block access ---------------------------------------------------------------------
    // Reads the header of the BGZF member at `off`; false at a clean end of file.
    bool block_at(int64_t off, Block &b) {
        if (off >= file_size) return false;
        if (off + 18 > file_size) {
            error = "truncated BGZF block header";
            return false;
        }
        const uint8_t *h = map + off;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) {
            error = "not a BGZF block (gzip member without the BC extra field)";
            return false;
        }
        const int xlen = rd16(h + 10);
        if (off + 12 + xlen > file_size) {
            error = "truncated BGZF extra field";
            return false;
        }
        // the BC subfield is normally first; walk the extra field to be safe
        const uint8_t *extra = h + 12;
        int bsize = -1;
        for (int i = 0; i + 4 <= xlen;) {
            const int slen = rd16(&extra[i + 2]);
            if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2 && i + 6 <= xlen) bsize = rd16(&extra[i + 4]);
            i += 4 + slen;
        }
        if (bsize < 0) {
            error = "BGZF block without BSIZE";
            return false;
        }
        b.file_off = off;
        b.csize = bsize + 1;
        if (off + b.csize > file_size || b.csize < 12 + xlen + 8) {
            error = "truncated BGZF block";
            return false;
        }
        b.usize = (int32_t)rd32(h + b.csize - 4);
        if (b.usize < 0 || b.usize > 65536) {     // a BGZF member holds at most 64 KB of data
            error = "corrupt BGZF block (ISIZE)";
            return false;
        }
        return true;
    }

    // Inflates the members described by `blocks` into out[i]; `threads` inflate in parallel.
    bool inflate_blocks(const std::vector<Block> &blocks, std::vector<std::vector<uint8_t>> &out) {
        if (blocks.empty()) return true;
        out.assign(blocks.size(), {});
        std::vector<int> ok(blocks.size(), 1);
        auto work = [&](size_t a, size_t b) {
            for (size_t i = a; i < b; ++i) {
                const Block &bl = blocks[i];
                const uint8_t *src = map + bl.file_off;
                const int xlen = rd16(src + 10);
                const int hdr = 12 + xlen;
                out[i].resize((size_t)bl.usize);
                if (bl.usize == 0) continue;
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) != Z_OK) { ok[i] = 0; continue; }
                zs.next_in = const_cast<Bytef *>(src + hdr);
                zs.avail_in = (uInt)(bl.csize - hdr - 8);
                zs.next_out = out[i].data();
                zs.avail_out = (uInt)bl.usize;
                const int rc = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (rc != Z_STREAM_END || zs.avail_out != 0) ok[i] = 0;
            }
        };
        const int nt = (int)std::min<size_t>((size_t)std::max(threads, 1), blocks.size());
        if (nt <= 1) {
            work(0, blocks.size());
        } else {
            std::vector<std::thread> pool;
            const size_t chunk = (blocks.size() + nt - 1) / nt;
            for (int t = 0; t < nt; ++t) {
                const size_t a = t * chunk, b = std::min(blocks.size(), a + chunk);
                if (a >= b) continue;
                try {
                    pool.emplace_back(work, a, b);
                } catch (const std::exception &) {
                    work(a, b);   // no thread to be had: the chunk is inflated here
                }
            }
            for (auto &th : pool) th.join();
        }
        for (int v : ok)
            if (!v) {
                error = "BGZF block does not inflate (corrupt file?)";
                return false;
            }
        return true;
    }
};

