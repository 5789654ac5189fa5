// Host-side limits of the file-to-file path (no GPU involved): how fast can N threads move a RAM-backed file into / out of
// 16 MiB buffers?  This is what ssd_run_files' I/O lanes do (csrc/ssd_b200.cu: upload_file_worker / download_file_worker).
// Reads: pread against memcpy from a mapping.  Writes of a NEW file (every page has to be allocated and zeroed by the
// kernel, the expensive part): pwrite, memcpy into a mapping, the same after fallocate (pages allocated in bulk, by one
// thread or by all threads on their own ranges), the same with MADV_POPULATE_WRITE per chunk; and writes over an EXISTING
// file (pages already there: what a rerun into the same output path sees when the file is not truncated first).
//   g++ -O2 -pthread -o /tmp/host_io_probe profiles/host_io_probe.cpp && /tmp/host_io_probe /dev/shm/probe.bin 3000000000
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const size_t kChunk = 16 << 20;

enum Mode { PREAD, MAPREAD, PWRITE, MAPWRITE, FALLOC1_MAPWRITE, FALLOCN_MAPWRITE, POPULATE_MAPWRITE, FALLOCN_PWRITE, OVERWRITE_MAP, OVERWRITE_PWRITE };

static double run(const char* path, size_t size, int nt, Mode mode)
{
    std::vector<char*> bufs(nt);
    for (int t = 0; t < nt; t++) {
        bufs[t] = (char*)aligned_alloc(4096, kChunk);
        memset(bufs[t], 65 + t, kChunk);
    }
    const bool write = mode >= PWRITE, over = mode >= OVERWRITE_MAP;
    const bool mapped = mode == MAPREAD || mode == MAPWRITE || mode == FALLOC1_MAPWRITE || mode == FALLOCN_MAPWRITE || mode == POPULATE_MAPWRITE || mode == OVERWRITE_MAP;
    if (write && !over) unlink(path);
    const int fd = open(path, write ? O_RDWR | O_CREAT : O_RDONLY, 0644);
    const double t0 = now();
    char* map = nullptr;
    if (write && !over && ftruncate(fd, (off_t)size)) return -1;
    if (mode == FALLOC1_MAPWRITE && fallocate(fd, 0, 0, (off_t)size)) return -2;
    if (mapped) map = (char*)mmap(nullptr, size, write ? PROT_READ | PROT_WRITE : PROT_READ, MAP_SHARED, fd, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            for (size_t k = t; k * kChunk < size; k += nt) {
                const size_t off = k * kChunk, want = std::min(kChunk, size - off);
                if (mode == FALLOCN_MAPWRITE || mode == FALLOCN_PWRITE) fallocate(fd, 0, (off_t)off, (off_t)want);
                if (mode == POPULATE_MAPWRITE) madvise(map + off, want, MADV_POPULATE_WRITE);
                if (mapped) {
                    if (write) memcpy(map + off, bufs[t], want);
                    else memcpy(bufs[t], map + off, want);
                } else {
                    size_t done = 0;
                    while (done < want) {
                        const ssize_t r = write ? pwrite(fd, bufs[t] + done, want - done, off + done) : pread(fd, bufs[t] + done, want - done, off + done);
                        if (r <= 0) break;
                        done += r;
                    }
                }
            }
        });
    for (auto& x : th) x.join();
    if (map) munmap(map, size);
    if (over && ftruncate(fd, (off_t)size)) return -3;
    close(fd);
    const double dt = now() - t0;
    for (char* b : bufs) free(b);
    return size / 1e9 / dt;
}

int main(int argc, char** argv)
{
    const char* path = argc > 1 ? argv[1] : "/dev/shm/ssd_io_probe.bin";
    const size_t size = argc > 2 ? strtoull(argv[2], nullptr, 10) : 3000000000ull;
    printf("%u hardware threads, %zu bytes, %s (GB/s)\n", std::thread::hardware_concurrency(), size, path);
    for (int nt : {4, 8, 16}) {
        const double w0 = run(path, size, nt, PWRITE);
        const double r0 = run(path, size, nt, PREAD);
        const double r1 = run(path, size, nt, MAPREAD);
        const double w1 = run(path, size, nt, MAPWRITE);
        const double w2 = run(path, size, nt, FALLOC1_MAPWRITE);
        const double w3 = run(path, size, nt, FALLOCN_MAPWRITE);
        const double w4 = run(path, size, nt, POPULATE_MAPWRITE);
        const double w5 = run(path, size, nt, FALLOCN_PWRITE);
        const double w6 = run(path, size, nt, OVERWRITE_MAP);
        const double w7 = run(path, size, nt, OVERWRITE_PWRITE);
        printf("threads %2d: read: pread %6.2f  mapped %6.2f | new file: pwrite %6.2f  mapped %6.2f  fallocate(1)+mapped %6.2f  fallocate(N)+mapped %6.2f  "
               "populate+mapped %6.2f  fallocate(N)+pwrite %6.2f | existing file: mapped %6.2f  pwrite %6.2f\n",
               nt, r0, r1, w0, w1, w2, w3, w4, w5, w6, w7);
    }
    unlink(path);
    return 0;
}
