// gs_multi.cpp — the native multi-GPU host driver: one host thread + one gs_ctx + its streams per GPU, frames sharded in
// contiguous blocks, no collective (frames are independent — Alfa_GS::process_pointcloud keeps no cross-frame
// state, src/alfa_gs.cpp:196-239 — so the reference's one-callback-per-cloud model, src/alfa_node.cpp:298-319, shards trivially).
//
// Everything here sits ABOVE the single-GPU C ABI (gs_create / gs_process_batch_host_strided): a worker thread is bound to the CPUs
// next to its GPU (the PCI device's local_cpulist) before it creates its context, so that the context's pinned staging and the
// copy submission stay on the GPU's NUMA node; gs_multi_host_alloc gives callers pinned batch buffers whose pages are first
// touched by the worker that will copy them (same block partition as the frames), which is what keeps eight concurrent H2D
// streams from sharing one memory controller.
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gs_b200.h"

namespace {

// CPUs local to a CUDA device, from sysfs (empty when unknown: the thread keeps its affinity)
std::vector<int> local_cpus(int device, int* numa_node) {
    std::vector<int> cpus;
    *numa_node = -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return cpus;
    }
    std::string id(bus);
    for (char& ch : id) ch = static_cast<char>(tolower(ch));
    const std::string base = "/sys/bus/pci/devices/" + id;
    if (FILE* f = fopen((base + "/numa_node").c_str(), "r")) {
        if (fscanf(f, "%d", numa_node) != 1) *numa_node = -1;
        fclose(f);
    }
    if (FILE* f = fopen((base + "/local_cpulist").c_str(), "r")) {
        char line[4096] = {0};
        if (fgets(line, sizeof(line), f)) {
            for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
                int a = 0, b = 0;
                const int n = sscanf(tok, "%d-%d", &a, &b);
                if (n == 1) b = a;
                if (n >= 1)
                    for (int c = a; c <= b; ++c) cpus.push_back(c);
            }
        }
        fclose(f);
    }
    return cpus;
}

struct Worker {
    int device = 0, numa_node = -1;
    bool bound = false;
    gs_ctx* ctx = nullptr;
    int create_rc = GS_OK;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = true, quit = false;
    int rc = GS_OK;
    std::string err;
};

}  // namespace

struct gs_multi {
    gs_params params;
    std::vector<Worker*> workers;
    std::string err;
};

namespace {

void worker_main(Worker* w, const gs_params* p) {
    // bind to the GPU's CPUs first: whatever this thread allocates and touches from here on lands on that node
    const std::vector<int> cpus = local_cpus(w->device, &w->numa_node);
    if (!cpus.empty()) {
        cpu_set_t allowed, want;
        CPU_ZERO(&want);
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
            int n = 0;
            for (int c : cpus)
                if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); ++n; }
            if (n > 0 && sched_setaffinity(0, sizeof(want), &want) == 0) w->bound = true;
        }
    }
    cudaSetDevice(w->device);
    w->create_rc = gs_create(p, w->device, &w->ctx);
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->done = true;
    }
    w->cv.notify_all();
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) break;
            job = std::move(w->job);
            w->has_job = false;
        }
        job();
        {
            std::lock_guard<std::mutex> lk(w->mu);
            w->done = true;
        }
        w->cv.notify_all();
    }
    if (w->ctx) gs_destroy(w->ctx);
    w->ctx = nullptr;
}

void submit(Worker* w, std::function<void()> job) {
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = std::move(job);
        w->has_job = true;
        w->done = false;
    }
    w->cv.notify_all();
}
void wait_done(Worker* w) {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
}

// contiguous block partition: shard g of G owns [g * n / G, (g + 1) * n / G)
inline long long block_begin(long long n, int g, int G) { return n * g / G; }

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int gs_multi_create(const gs_params* p, const int* devices, int n_devices, gs_multi** out) {
    if (!p || !out || n_devices < 0) return GS_EINVAL;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
        cudaGetLastError();
        return GS_ECUDA;  // no CPU fallback
    }
    gs_multi* m = new gs_multi;
    m->params = *p;
    const int n = (devices && n_devices > 0) ? n_devices : visible;
    for (int g = 0; g < n; ++g) {
        const int dev = (devices && n_devices > 0) ? devices[g] : g;
        if (dev < 0 || dev >= visible) {
            for (Worker* w : m->workers) {
                { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
                w->cv.notify_all();
                w->th.join();
                delete w;
            }
            delete m;
            return GS_EINVAL;
        }
        Worker* w = new Worker;
        w->device = dev;
        w->done = false;
        w->th = std::thread(worker_main, w, &m->params);
        m->workers.push_back(w);
    }
    int rc = GS_OK;
    for (Worker* w : m->workers) {
        wait_done(w);
        if (w->create_rc != GS_OK) rc = w->create_rc;
    }
    if (rc != GS_OK) {
        for (Worker* w : m->workers) {
            { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
            w->cv.notify_all();
            w->th.join();
            delete w;
        }
        delete m;
        return rc;
    }
    *out = m;
    return GS_OK;
}

void gs_multi_destroy(gs_multi* m) {
    if (!m) return;
    for (Worker* w : m->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        w->th.join();
        delete w;
    }
    delete m;
}

int gs_multi_n_gpus(const gs_multi* m) { return m ? static_cast<int>(m->workers.size()) : 0; }

int gs_multi_worker_info(const gs_multi* m, int shard, int* device, int* numa_node, int* cpu_bound) {
    if (!m || shard < 0 || shard >= static_cast<int>(m->workers.size())) return GS_EINVAL;
    const Worker* w = m->workers[shard];
    if (device) *device = w->device;
    if (numa_node) *numa_node = w->numa_node;
    if (cpu_bound) *cpu_bound = w->bound ? 1 : 0;
    return GS_OK;
}

const char* gs_multi_last_error(const gs_multi* m) { return m ? m->err.c_str() : "null handle"; }

// Pinned host memory for batch buffers, `bytes` long: shard g's worker first-touches the g-th block of it (block partition by
// bytes, which is where shard g's frames lie when frames are of similar size), then the whole range is page-locked for all devices.
int gs_multi_host_alloc(gs_multi* m, size_t bytes, void** out) {
    if (!m || !out || bytes == 0) return GS_EINVAL;
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return GS_ENOMEM;
    const int G = static_cast<int>(m->workers.size());
    for (int g = 0; g < G; ++g) {
        const size_t b0 = static_cast<size_t>(block_begin(static_cast<long long>(bytes), g, G));
        const size_t b1 = static_cast<size_t>(block_begin(static_cast<long long>(bytes), g + 1, G));
        unsigned char* base = static_cast<unsigned char*>(p);
        submit(m->workers[g], [base, b0, b1] {
            for (size_t i = b0; i < b1; i += 4096) base[i] = 0;  // first touch decides the NUMA node of the page
        });
    }
    for (Worker* w : m->workers) wait_done(w);
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(p, bytes);
        return GS_ENOMEM;
    }
    *out = p;
    return GS_OK;
}

int gs_multi_host_free(gs_multi* m, void* p, size_t bytes) {
    if (!m || !p) return GS_EINVAL;
    cudaHostUnregister(p);
    munmap(p, bytes);
    return GS_OK;
}

// The host batch call of one GPU (gs_process_batch_host_strided) on every GPU at once: shard g processes the frames
// [g * F / G, (g + 1) * F / G).  h_no_ground_out: n_frames * P u16 (may be NULL).  Clouds: shard g packs its frames back to back
// starting at record g * xyz4_capacity / G of h_xyz4_out (the shards do not know each other's sizes); h_out_first[f] /
// h_out_count[f] give the first record and the number of records of frame f.  Synchronous.
int gs_multi_process_batch_host(gs_multi* m, const void* h_points, size_t stride_bytes, const int64_t* h_offsets, int n_frames,
                                uint16_t* h_no_ground_out, float* h_xyz4_out, uint64_t xyz4_capacity, int64_t* h_out_first,
                                uint32_t* h_out_count) {
    if (!m || !h_offsets || n_frames < 0 || (n_frames && !h_points)) return GS_EINVAL;
    if (h_xyz4_out && (!h_out_first || !h_out_count)) return GS_EINVAL;
    if (n_frames == 0) return GS_OK;
    gs_derived d;
    if (gs_derive(&m->params, &d) != GS_OK) return GS_EINVAL;
    const size_t P = static_cast<size_t>(d.n_rows) * d.n_cols;
    const int G = static_cast<int>(m->workers.size());
    std::vector<std::vector<int64_t>> shard_offs(G);
    for (int g = 0; g < G; ++g) {
        Worker* w = m->workers[g];
        const int f0 = static_cast<int>(block_begin(n_frames, g, G)), f1 = static_cast<int>(block_begin(n_frames, g + 1, G));
        const uint64_t c0 = xyz4_capacity / G * g, cap = xyz4_capacity / G;
        shard_offs[g].assign(static_cast<size_t>(f1 - f0) + 1, 0);
        int64_t* so = shard_offs[g].data();
        submit(w, [=] {
            w->rc = GS_OK;
            if (f1 == f0) return;
            w->rc = gs_process_batch_host_strided(w->ctx, h_points, stride_bytes, h_offsets + f0, f1 - f0,
                                                  h_no_ground_out ? h_no_ground_out + static_cast<size_t>(f0) * P : nullptr,
                                                  h_xyz4_out ? h_xyz4_out + 4 * c0 : nullptr, cap, so);
            if (w->rc != GS_OK) w->err = gs_last_error(w->ctx);
            else if (h_out_first)
                for (int f = f0; f < f1; ++f) {
                    h_out_first[f] = static_cast<int64_t>(c0) + so[f - f0];
                    h_out_count[f] = static_cast<uint32_t>(so[f - f0 + 1] - so[f - f0]);
                }
        });
    }
    int rc = GS_OK;
    for (Worker* w : m->workers) {
        wait_done(w);
        if (w->rc != GS_OK && rc == GS_OK) {
            rc = w->rc;
            m->err = "GPU " + std::to_string(w->device) + ": " + w->err;
        }
    }
    return rc;
}

}  // extern "C"
#pragma GCC visibility pop
