// Task counter shared by the ranks of one box (include/ncb.h: ncb_counter_*): the dynamic half of the
// transcript sharding.  The reference distributes transcripts through a multiprocessing queue
// (run.py:104-131); here every rank knows the task list and only the cursor is shared -- a 64-bit word in
// POSIX shared memory, advanced with an atomic fetch-add.  Host code, no CUDA calls.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <new>
#include <string>

#include "ncb.h"

struct ncb_counter {
    int64_t *word = nullptr;
    std::string name;
};

extern "C" {

int ncb_counter_open(const char *name, int create, ncb_counter **out) {
    if (!out) return NCB_ERR_INVALID;
    *out = nullptr;
    if (!name || name[0] != '/' || strlen(name) < 2 || strlen(name) > 200) return NCB_ERR_INVALID;
    const int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return NCB_ERR_INVALID;
    if (create && ftruncate(fd, 64) != 0) { close(fd); return NCB_ERR_NOMEM; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < (off_t)sizeof(int64_t)) { close(fd); return NCB_ERR_INVALID; }
    void *p = mmap(nullptr, 64, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return NCB_ERR_NOMEM;
    ncb_counter *c = new (std::nothrow) ncb_counter();
    if (!c) { munmap(p, 64); return NCB_ERR_NOMEM; }
    c->word = (int64_t *)p;
    try {
        c->name = name;
    } catch (const std::exception &) {
        munmap(p, 64);
        delete c;
        return NCB_ERR_NOMEM;
    }
    if (create) __atomic_store_n(c->word, (int64_t)0, __ATOMIC_SEQ_CST);
    *out = c;
    return NCB_OK;
}

int64_t ncb_counter_add(ncb_counter *c, int64_t n) {
    if (!c) return -1;
    return __atomic_fetch_add(c->word, n, __ATOMIC_SEQ_CST);
}

void ncb_counter_set(ncb_counter *c, int64_t value) {
    if (c) __atomic_store_n(c->word, value, __ATOMIC_SEQ_CST);
}

void ncb_counter_close(ncb_counter *c, int unlink) {
    if (!c) return;
    munmap(c->word, 64);
    if (unlink) shm_unlink(c->name.c_str());
    delete c;
}

}  // extern "C"
