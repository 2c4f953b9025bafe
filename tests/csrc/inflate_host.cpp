// Host build of nanocompore_b200/csrc/inflate_core.cuh (the DEFLATE decoder of the device-side BAM
// reader) for the CPU test-suite (tests/test_inflate_host.py): the same functions the CUDA warp runs,
// driven by a one-lane "warp", checked against zlib.  Test infrastructure only.
#include "inflate_core.cuh"
#include <string.h>
#include <vector>

extern "C" {

// raw DEFLATE payload -> out[usize]; returns an ncb_inflate error code (0 = ok)
int ncb_test_inflate(const uint8_t *payload, long long cbytes, uint8_t *out, long long usize) {
    // the device gets 4-byte aligned payloads followed by zero padding; reading past it returns 0
    const uint32_t n_words = (uint32_t)((cbytes + 3) / 4);
    std::vector<uint32_t> words(n_words + 2, 0u);
    if (cbytes > 0) memcpy(words.data(), payload, (size_t)cbytes);
    static ncb_inflate::Tables tables;      // 4 KB: what a warp keeps in shared memory
    ncb_inflate::SerialWarp w;
    return ncb_inflate::inflate_member(w, tables, words.data(), n_words, cbytes, out, usize);
}

int ncb_test_inflate_tables_bytes(void) { return (int)sizeof(ncb_inflate::Tables); }

}
