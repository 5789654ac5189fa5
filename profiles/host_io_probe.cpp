// Host-side limits of the file-to-file path (no GPU involved): how fast can N threads move a RAM-backed file into / out of
// 16 MiB buffers with pread / pwrite against memcpy from / into a mapping?  This is what ssd_run_files' I/O lanes do
// (csrc/ssd_b200.cu: upload_file_worker / download_file_worker).
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

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const size_t kChunk = 16 << 20;

static double run(const char* path, size_t size, int nt, bool write, bool mapped)
{
    std::vector<char*> bufs(nt);
    for (int t = 0; t < nt; t++) {
        bufs[t] = (char*)aligned_alloc(4096, kChunk);
        memset(bufs[t], 65 + t, kChunk);
    }
    if (write) unlink(path);
    const int fd = open(path, write ? O_RDWR | O_CREAT | O_TRUNC : O_RDONLY, 0644);
    const double t0 = now();
    char* map = nullptr;
    if (mapped) {
        if (write && ftruncate(fd, (off_t)size)) return -1;
        map = (char*)mmap(nullptr, size, write ? PROT_READ | PROT_WRITE : PROT_READ, MAP_SHARED, fd, 0);
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            for (size_t k = t; k * kChunk < size; k += nt) {
                const size_t off = k * kChunk, want = std::min(kChunk, size - off);
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
    close(fd);
    const double dt = now() - t0;
    for (char* b : bufs) free(b);
    return size / 1e9 / dt;
}

int main(int argc, char** argv)
{
    const char* path = argc > 1 ? argv[1] : "/dev/shm/ssd_io_probe.bin";
    const size_t size = argc > 2 ? strtoull(argv[2], nullptr, 10) : 3000000000ull;
    printf("%u hardware threads, %zu bytes, %s\n", std::thread::hardware_concurrency(), size, path);
    for (int nt : {4, 8, 16}) {
        const double w0 = run(path, size, nt, true, false);
        const double r0 = run(path, size, nt, false, false);
        const double r1 = run(path, size, nt, false, true);
        const double w1 = run(path, size, nt, true, true);
        printf("threads %2d: pread %6.2f GB/s  mmap+memcpy read %6.2f GB/s | pwrite (new file) %6.2f GB/s  mmap+memcpy write (new file) %6.2f GB/s\n", nt, r0, r1, w0, w1);
    }
    unlink(path);
    return 0;
}
