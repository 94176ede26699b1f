// d2h_probe — what the BOX can move between its GPUs and host memory, independent of the solver.
//
// One host thread per GPU.  For each placement policy of the pinned host buffers the probe times (a) every GPU alone
// and (b) all GPUs at once, device->host and host->device, with plain cudaMemcpyAsync on one stream per GPU:
//   default : cudaHostAlloc from the main thread (what torch's pin_memory() gives every rank: first-touch on whatever
//             node the allocating thread runs on)
//   local   : each GPU's thread first binds itself to the CPUs of the GPU's NUMA node (sysfs numa_node / cpulist), then
//             cudaHostAlloc's and touches its buffer
//   mbind   : mmap + mbind(MPOL_BIND, node of the GPU) + touch + cudaHostRegister
// The concurrent device->host figure is the ceiling the host-buffer path of mcf_runmicro (80 B per cell-hour of FP64
// results) can reach at N GPUs; bench.py prints it next to e2e.pcie_gb_per_s.
//
// build: nvcc -O2 -std=c++17 -o d2h_probe d2h_probe.cu -lpthread      usage: d2h_probe [MiB per GPU = 1024] [reps = 4]
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Gpu {
    int dev = 0;
    char bus[32] = {0};
    int node = -1;
    std::vector<int> cpus;
};

static std::vector<int> parse_cpulist(const std::string& s) {
    std::vector<int> out;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        if (tok.empty()) continue;
        int a = 0, b = 0;
        if (sscanf(tok.c_str(), "%d-%d", &a, &b) == 2) {
            for (int i = a; i <= b; ++i) out.push_back(i);
        } else if (sscanf(tok.c_str(), "%d", &a) == 1) out.push_back(a);
    }
    return out;
}
static std::string slurp(const std::string& path) {
    std::ifstream f(path);
    std::string s;
    std::getline(f, s);
    return s;
}
static bool bind_cpus(const std::vector<int>& cpus) {
    if (cpus.empty()) return false;
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed);
    sched_getaffinity(0, sizeof allowed, &allowed);
    CPU_ZERO(&want);
    int n = 0;
    for (int c : cpus)
        if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) {
            CPU_SET(c, &want);
            ++n;
        }
    if (!n) return false;
    return pthread_setaffinity_np(pthread_self(), sizeof want, &want) == 0;
}
static long mbind_node(void* p, size_t bytes, int node) {
    if (node < 0) return -1;
    unsigned long mask[16] = {0};
    mask[node / 64] |= 1UL << (node % 64);
    return syscall(SYS_mbind, p, bytes, 2 /* MPOL_BIND */, mask, 1024UL, 0UL);
}

struct Barrier {
    std::atomic<int> count{0}, gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        const int g = gen.load();
        if (count.fetch_add(1) + 1 == n) {
            count.store(0);
            gen.fetch_add(1);
        } else
            while (gen.load() == g) sched_yield();
    }
};

int main(int argc, char** argv) {
    const size_t mib = argc > 1 ? (size_t)atoll(argv[1]) : 1024;
    const int reps = argc > 2 ? atoi(argv[2]) : 4;
    const size_t bytes = mib << 20;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        printf("{\"error\": \"no CUDA device\"}\n");
        return 1;
    }
    std::vector<Gpu> gpus(ndev);
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    sched_getaffinity(0, sizeof allowed, &allowed);
    fprintf(stderr, "host: %d CPUs allowed of %ld online; NUMA nodes online: %s\n", CPU_COUNT(&allowed),
            sysconf(_SC_NPROCESSORS_ONLN), slurp("/sys/devices/system/node/online").c_str());
    for (int d = 0; d < ndev; ++d) {
        gpus[d].dev = d;
        cudaDeviceGetPCIBusId(gpus[d].bus, sizeof gpus[d].bus, d);
        std::string b = gpus[d].bus;
        for (auto& ch : b) ch = (char)tolower(ch);
        const std::string nn = slurp("/sys/bus/pci/devices/" + b + "/numa_node");
        gpus[d].node = nn.empty() ? -1 : atoi(nn.c_str());
        std::string cl = slurp("/sys/bus/pci/devices/" + b + "/local_cpulist");
        gpus[d].cpus = parse_cpulist(cl);
        int usable = 0;
        for (int c : gpus[d].cpus) usable += (c < CPU_SETSIZE && CPU_ISSET(c, &allowed));
        fprintf(stderr, "gpu %d %s numa_node %d local_cpulist %s (%d of them allowed)\n", d, gpus[d].bus, gpus[d].node,
                cl.c_str(), usable);
    }
    std::vector<void*> dbuf(ndev, nullptr);
    std::vector<cudaStream_t> st(ndev);
    for (int d = 0; d < ndev; ++d) {
        cudaSetDevice(d);
        cudaMalloc(&dbuf[d], bytes);
        cudaMemset(dbuf[d], 1, bytes);
        cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking);
    }
    printf("{\"mib_per_gpu\": %zu, \"reps\": %d, \"n_gpus\": %d, \"cpus_allowed\": %d, \"policies\": {", mib, reps, ndev,
           CPU_COUNT(&allowed));
    const char* pol_names[3] = {"default", "local", "mbind"};
    for (int pol = 0; pol < 3; ++pol) {
        std::vector<void*> hbuf(ndev, nullptr);
        std::vector<int> ok(ndev, 1);
        std::vector<long> mb(ndev, 0);
        if (pol == 0) {
            for (int d = 0; d < ndev; ++d) {
                cudaSetDevice(d);
                if (cudaHostAlloc(&hbuf[d], bytes, cudaHostAllocDefault) != cudaSuccess) ok[d] = 0;
                else memset(hbuf[d], 0, bytes);
            }
        }
        // alone[dir][d], together[dir]
        std::vector<double> alone_d2h(ndev, 0), alone_h2d(ndev, 0);
        double all_d2h = 0, all_h2d = 0, all_bidir = 0;
        Barrier bar(ndev);
        std::vector<double> t0s(ndev), t1s(ndev);
        std::vector<std::thread> th;
        for (int d = 0; d < ndev; ++d) {
            th.emplace_back([&, d]() {
                cudaSetDevice(d);
                if (pol >= 1) bind_cpus(gpus[d].cpus);
                if (pol == 1) {
                    if (cudaHostAlloc(&hbuf[d], bytes, cudaHostAllocDefault) != cudaSuccess) ok[d] = 0;
                    else memset(hbuf[d], 0, bytes);
                } else if (pol == 2) {
                    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                    if (p == MAP_FAILED) ok[d] = 0;
                    else {
                        mb[d] = mbind_node(p, bytes, gpus[d].node);
                        memset(p, 0, bytes);
                        if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
                            ok[d] = 0;
                            munmap(p, bytes);
                        } else hbuf[d] = p;
                    }
                }
                bar.wait();
                auto run = [&](int dir) { // 0 d2h, 1 h2d, 2 both (two halves)
                    for (int r = 0; r < reps; ++r) {
                        if (dir == 0) cudaMemcpyAsync(hbuf[d], dbuf[d], bytes, cudaMemcpyDeviceToHost, st[d]);
                        else if (dir == 1) cudaMemcpyAsync(dbuf[d], hbuf[d], bytes, cudaMemcpyHostToDevice, st[d]);
                        else cudaMemcpyAsync(hbuf[d], dbuf[d], bytes, cudaMemcpyDeviceToHost, st[d]);
                    }
                    cudaStreamSynchronize(st[d]);
                };
                if (ok[d]) run(0); // warm
                // each GPU alone, in turn
                for (int dir = 0; dir < 2; ++dir)
                    for (int who = 0; who < ndev; ++who) {
                        bar.wait();
                        if (who == d && ok[d]) {
                            const double t0 = now();
                            run(dir);
                            const double gb = (double)bytes * reps / (now() - t0) / 1e9;
                            (dir ? alone_h2d : alone_d2h)[d] = gb;
                        }
                        bar.wait();
                    }
                // all together
                for (int dir = 0; dir < 2; ++dir) {
                    bar.wait();
                    t0s[d] = now();
                    if (ok[d]) run(dir);
                    t1s[d] = now();
                    bar.wait();
                    if (d == 0) {
                        const double t0 = *std::min_element(t0s.begin(), t0s.end());
                        const double t1 = *std::max_element(t1s.begin(), t1s.end());
                        int nok = 0;
                        for (int k : ok) nok += k;
                        (dir ? all_h2d : all_d2h) = (double)bytes * reps * nok / (t1 - t0) / 1e9;
                    }
                    bar.wait();
                }
                (void)all_bidir;
            });
        }
        for (auto& t : th) t.join();
        printf("%s\"%s\": {\"d2h_alone_gbs\": [", pol ? ", " : "", pol_names[pol]);
        for (int d = 0; d < ndev; ++d) printf("%s%.1f", d ? ", " : "", alone_d2h[d]);
        printf("], \"h2d_alone_gbs\": [");
        for (int d = 0; d < ndev; ++d) printf("%s%.1f", d ? ", " : "", alone_h2d[d]);
        printf("], \"d2h_all_gbs\": %.1f, \"h2d_all_gbs\": %.1f, \"ok\": [", all_d2h, all_h2d);
        for (int d = 0; d < ndev; ++d) printf("%s%d", d ? ", " : "", ok[d]);
        printf("], \"mbind_rc\": [");
        for (int d = 0; d < ndev; ++d) printf("%s%ld", d ? ", " : "", mb[d]);
        printf("]}");
        fflush(stdout);
        for (int d = 0; d < ndev; ++d) {
            if (!hbuf[d]) continue;
            if (pol == 2) {
                cudaHostUnregister(hbuf[d]);
                munmap(hbuf[d], bytes);
            } else cudaFreeHost(hbuf[d]);
        }
    }
    printf("}}\n");
    return 0;
}
