// AddressSanitizer / UBSan harness of the native BAM reader (host code of libncb, no CUDA):
// opens every file of a directory and reads all its references.
//
//   python tests/fuzz_bam.py 77 3000 --keep build/asan/files        # damaged BAMs, kept on disk
//   g++ -x c++ -std=c++17 -O1 -g -fsanitize=address,undefined -I include -I nanocompore_b200/csrc \
//       nanocompore_b200/csrc/reader.cu tests/csrc/asan_bam_harness.cpp -lz -lpthread -o build/asan/harness
//   ./build/asan/harness build/asan/files
//
// Round 1: 3 000 damaged files clean (one signed overflow on a corrupt l_seq found and fixed).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <dirent.h>
#include <string>
#include "ncb.h"
int main(int argc, char **argv) {
    DIR *d = opendir(argv[1]); struct dirent *e; int ok = 0, err = 0;
    while ((e = readdir(d))) {
        if (e->d_name[0] == '.') continue;
        std::string p = std::string(argv[1]) + "/" + e->d_name;
        ncb_bam *b = nullptr;
        int rc = ncb_bam_open(&b, p.c_str(), 2);
        if (rc != NCB_OK) { if (b) ncb_bam_close(b); ++err; continue; }
        bool bad = false;
        for (int r = 0; r < ncb_bam_n_references(b); ++r) {
            int64_t len = ncb_bam_reference_length(b, r), cap = ncb_bam_reference_records(b, r);
            if (len < 0 || len > 100000 || cap < 0 || cap > 100000) { bad = true; continue; }
            std::vector<float> I((size_t)(cap * len) + 1), D((size_t)(cap * len) + 1);
            std::vector<char> names((size_t)(cap + 1) * 256);
            int64_t n = ncb_bam_read_uncalled4(b, r, len, 5, 4, I.data(), D.data(), names.data(), 256, cap);
            if (n < 0) bad = true;
        }
        ncb_bam_close(b);
        if (bad) ++err; else ++ok;
    }
    closedir(d);
    printf("ok %d err %d\n", ok, err);
    return 0;
}
