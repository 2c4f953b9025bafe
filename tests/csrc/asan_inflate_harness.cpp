// AddressSanitizer / UBSan + mutation harness of the DEFLATE decoder of the device-side BAM reader
// (nanocompore_b200/csrc/inflate_core.cuh, the functions the CUDA warp runs): streams written by zlib at
// every level / strategy over signal-like, repetitive, random and empty data, intact and damaged (bit
// flips, byte changes, truncation, wrong ISIZE).  For every stream the decoder must agree with zlib's
// own inflate -- the acceptance rule of the host reader, reader.cu: Z_STREAM_END with exactly ISIZE
// bytes -- on accept / reject and, when accepted, on every byte.
//
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -I nanocompore_b200/csrc \
//       tests/csrc/asan_inflate_harness.cpp -lz -o build/asan/inflate && ./build/asan/inflate [streams]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <random>
#include <vector>
#include "inflate_core.cuh"

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t> &data, int level, int strategy) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
    std::vector<uint8_t> out(deflateBound(&zs, (uLong)data.size()) + 16);
    zs.next_in = const_cast<Bytef *>(data.data());
    zs.avail_in = (uInt)data.size();
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    deflate(&zs, Z_FINISH);
    out.resize(out.size() - zs.avail_out);
    deflateEnd(&zs);
    return out;
}

// the host reader's rule (reader.cu: inflate_blocks)
static bool zlib_accepts(const std::vector<uint8_t> &comp, size_t usize, std::vector<uint8_t> &out) {
    out.assign(usize + 1, 0);
    if (usize == 0) {   // the host reader does not inflate members with ISIZE 0
        out.resize(0);
        return true;
    }
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef *>(comp.data());
    zs.avail_in = (uInt)comp.size();
    zs.next_out = out.data();
    zs.avail_out = (uInt)usize;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
    inflateEnd(&zs);
    out.resize(usize);
    return ok;
}

static int ours(const std::vector<uint8_t> &comp, size_t usize, std::vector<uint8_t> &out) {
    const uint32_t n_words = (uint32_t)((comp.size() + 3) / 4);
    // exactly the words the device is given: no slack behind them (ASan sees any read past the padding)
    std::vector<uint32_t> words(n_words + 2, 0u);
    if (!comp.empty()) memcpy(words.data(), comp.data(), comp.size());
    out.assign(usize, 0);
    static ncb_inflate::Tables tables;
    ncb_inflate::SerialWarp w;
    return ncb_inflate::inflate_member(w, tables, words.data(), n_words, (int64_t)comp.size(), out.data(), (int64_t)usize);
}

int main(int argc, char **argv) {
    const int streams = argc > 1 ? atoi(argv[1]) : 400;
    std::mt19937 g(11);
    long checked = 0, accepted = 0, rejected = 0;
    for (int s = 0; s < streams; ++s) {
        const size_t n = s % 7 == 0 ? g() % 64 : (s % 5 == 0 ? 65536 - g() % 300 : g() % 20000);
        std::vector<uint8_t> data(n);
        switch (s % 4) {
            case 0: for (size_t i = 0; i + 1 < n; i += 2) { const int v = (int)(g() % 6000) - 3000; data[i] = (uint8_t)v; data[i + 1] = (uint8_t)(v >> 8); } break;
            case 1: for (size_t i = 0; i < n; ++i) data[i] = (uint8_t)("nanocompore"[i % (3 + s % 9)]); break;
            case 2: for (size_t i = 0; i < n; ++i) data[i] = (uint8_t)g(); break;
            default: for (size_t i = 0; i < n; ++i) data[i] = (uint8_t)(g() % 3); break;
        }
        const int level = (int)(g() % 10);
        const int strategies[5] = {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED};
        const std::vector<uint8_t> comp = deflate_raw(data, level, strategies[g() % 5]);
        for (int m = 0; m < 40; ++m) {
            std::vector<uint8_t> c = comp;
            size_t usize = n;
            if (m > 0) {
                const int kind = (int)(g() % 6);
                const int edits = 1 + (int)(g() % 3);
                for (int e = 0; e < edits && !c.empty(); ++e) {
                    const size_t at = (m % 3 == 0) ? g() % (c.size() < 40 ? c.size() : 40) : g() % c.size();   // headers get their share
                    if (kind == 0) c[at] ^= (uint8_t)(1u << (g() % 8));
                    else if (kind == 1) c[at] = (uint8_t)g();
                    else if (kind == 2) c.resize(at);
                    else if (kind == 3) { const long u = (long)n + (long)(g() % 9) - 4; usize = (size_t)(u < 0 ? 0 : u); }
                    else if (kind == 4) c.insert(c.begin() + at, (uint8_t)g());
                    else c.push_back((uint8_t)g());          // trailing bytes after the final block are ignored
                }
            }
            std::vector<uint8_t> want, got;
            const bool ok_z = zlib_accepts(c, usize, want);
            const int rc = usize == 0 ? 0 : ours(c, usize, got);      // ISIZE 0: not inflated (decode.cu skips such members too)
            if (usize == 0) got.clear();
            ++checked;
            if (ok_z != (rc == 0)) {
                printf("MISMATCH stream %d mutation %d: zlib %s, decoder rc %d (n %zu, usize %zu, csize %zu)\n", s, m,
                       ok_z ? "accepts" : "rejects", rc, n, usize, c.size());
                printf("payload:");
                for (size_t i = 0; i < c.size() && i < 64; ++i) printf(" %02x", c[i]);
                printf("\n");
                return 1;
            }
            if (ok_z) {
                ++accepted;
                if (got != want) {
                    printf("BYTES DIFFER stream %d mutation %d\n", s, m);
                    return 1;
                }
            } else {
                ++rejected;
            }
        }
    }
    printf("inflate harness: %ld streams checked against zlib (%ld accepted, %ld rejected), all agree\n", checked, accepted, rejected);
    return 0;
}
