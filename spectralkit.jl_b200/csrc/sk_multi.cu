// Multi-GPU entry points that work on DEVICE-resident shards (SURVEY.md §8(b) "sk_*_sharded(comm, …)", §8(e)):
// one process drives G devices; points are independent, so every device evaluates its own contiguous slice with the
// ordinary kernels; the only communication is (C1) one broadcast of the coefficients and (C2) an all-gather of the
// result slices when the caller asks for a device-resident full output.  NCCL is bound at run time (dlopen of
// libnccl.so.2 — the copy already loaded by the host program if there is one), so the library itself has no NCCL
// link dependency and single-GPU users never touch it.  The reference is single-threaded and has no counterpart.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "sk_launch.h"

using skh::set_error;

namespace {

// the few NCCL 2.x entry points this file needs (ABI-stable since 2.0)
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0, ncclFloat64_ = 8 };
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

const NcclApi &nccl() {
    static const NcclApi api = [] {
        NcclApi a;
        const char *names[] = {getenv("SK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.h) break;
            a.why = dlerror();
        }
        if (!a.h) return a;
#define SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.h, name))
        SYM(GetVersion, "ncclGetVersion"); SYM(CommInitAll, "ncclCommInitAll"); SYM(CommDestroy, "ncclCommDestroy");
        SYM(AllGather, "ncclAllGather"); SYM(Broadcast, "ncclBroadcast"); SYM(GroupStart, "ncclGroupStart");
        SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
        if (!a.CommInitAll || !a.CommDestroy || !a.AllGather || !a.Broadcast || !a.GroupStart || !a.GroupEnd) {
            a.why = "libnccl lacks a required symbol"; dlclose(a.h); a.h = nullptr;
        }
        return a;
    }();
    return api;
}

#define NK(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess_) {                                                                         \
            set_error(std::string(#call) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r_) : "NCCL error")); \
            return SK_ERR_NCCL;                                                                           \
        }                                                                                                 \
    } while (0)
#define CKM(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) { set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return SK_ERR_CUDA; } \
    } while (0)

}  // namespace

struct sk_comm {
    int ngpu = 0;
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;      // one non-blocking stream per device: kernels and collectives of a call
    std::vector<cudaEvent_t> ev0, ev1;
    // gather mode 2 (result slices pushed to the peers over NVLink while the next chunk is computed): every pair of devices
    // has peer access, and device g has one copy stream per destination h
    bool peer_ok = false;
    std::vector<std::vector<cudaStream_t>> cstreams;      // [g][h]
};

extern "C" {

int sk_comm_create(int ngpu, const int *devices, sk_comm **out) {
    if (!out) SK_FAIL(SK_ERR_ARGUMENT, "out is NULL");
    *out = nullptr;
    if (ngpu < 1 || !devices) SK_FAIL(SK_ERR_ARGUMENT, "need at least one device");
    int ndev = 0;
    int rc = sk_device_count(&ndev);
    if (rc != SK_OK) return rc;
    for (int g = 0; g < ngpu; g++) {
        if (devices[g] < 0 || devices[g] >= ndev) SK_FAIL(SK_ERR_ARGUMENT, "device index out of range");
        for (int h = 0; h < g; h++) if (devices[h] == devices[g]) SK_FAIL(SK_ERR_ARGUMENT, "devices must be distinct");
    }
    const NcclApi &n = nccl();
    if (!n.h) SK_FAIL(SK_ERR_NCCL, "NCCL is not available: " + n.why);
    sk_comm *c = new (std::nothrow) sk_comm();
    if (!c) SK_FAIL(SK_ERR_ARGUMENT, "out of memory");
    c->ngpu = ngpu;
    c->devices.assign(devices, devices + ngpu);
    c->comms.assign((size_t)ngpu, nullptr);
    ncclResult_t r = n.CommInitAll(c->comms.data(), ngpu, devices);
    if (r != ncclSuccess_) {
        set_error(std::string("ncclCommInitAll: ") + (n.GetErrorString ? n.GetErrorString(r) : "NCCL error"));
        delete c;
        return SK_ERR_NCCL;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    c->streams.assign((size_t)ngpu, nullptr); c->ev0.assign((size_t)ngpu, nullptr); c->ev1.assign((size_t)ngpu, nullptr);
    for (int g = 0; g < ngpu; g++) {
        cudaError_t e = cudaSetDevice(devices[g]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->streams[(size_t)g], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&c->ev0[(size_t)g]);
        if (e == cudaSuccess) e = cudaEventCreate(&c->ev1[(size_t)g]);
        if (e != cudaSuccess) {
            set_error(std::string("sk_comm_create: ") + cudaGetErrorString(e));
            cudaSetDevice(prev);
            sk_comm_destroy(c);
            return SK_ERR_CUDA;
        }
    }
    // peer access between every pair (NVLink / NVSwitch): needed by the overlapped gather; without it mode 2 falls back to NCCL
    c->peer_ok = ngpu > 1;
    for (int g = 0; g < ngpu && c->peer_ok; g++)
        for (int h = 0; h < ngpu && c->peer_ok; h++) {
            if (h == g) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[g], devices[h]) != cudaSuccess || !can) c->peer_ok = false;
        }
    if (c->peer_ok) {
        c->cstreams.assign((size_t)ngpu, std::vector<cudaStream_t>((size_t)ngpu, nullptr));
        for (int g = 0; g < ngpu && c->peer_ok; g++) {
            cudaSetDevice(devices[g]);
            for (int h = 0; h < ngpu && c->peer_ok; h++) {
                if (h == g) continue;
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[h], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
                if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->cstreams[(size_t)g][(size_t)h], cudaStreamNonBlocking);
                if (e != cudaSuccess) { (void)cudaGetLastError(); c->peer_ok = false; }
            }
        }
    }
    cudaSetDevice(prev);
    *out = c;
    return SK_OK;
}

int sk_comm_destroy(sk_comm *c) {
    if (!c) return SK_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    for (int g = 0; g < c->ngpu; g++) {
        cudaSetDevice(c->devices[(size_t)g]);
        if ((size_t)g < c->streams.size() && c->streams[(size_t)g]) { cudaStreamSynchronize(c->streams[(size_t)g]); cudaStreamDestroy(c->streams[(size_t)g]); }
        if ((size_t)g < c->cstreams.size())
            for (cudaStream_t s : c->cstreams[(size_t)g]) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
        if ((size_t)g < c->ev0.size() && c->ev0[(size_t)g]) cudaEventDestroy(c->ev0[(size_t)g]);
        if ((size_t)g < c->ev1.size() && c->ev1[(size_t)g]) cudaEventDestroy(c->ev1[(size_t)g]);
        if (c->comms[(size_t)g] && nccl().CommDestroy) nccl().CommDestroy(c->comms[(size_t)g]);
    }
    cudaSetDevice(prev);
    delete c;
    return SK_OK;
}

int sk_comm_size(const sk_comm *c, int *ngpu, int *nccl_version) {
    if (!c) SK_FAIL(SK_ERR_ARGUMENT, "comm is NULL");
    if (ngpu) *ngpu = c->ngpu;
    if (nccl_version) { *nccl_version = 0; if (nccl().GetVersion) nccl().GetVersion(nccl_version); }
    return SK_OK;
}

// C1: the coefficients live on device `root`; every other device receives a copy (count doubles).
int sk_broadcast_coefficients(sk_comm *c, double *const *theta_dev, int64_t count, int root) {
    if (!c || !theta_dev) SK_FAIL(SK_ERR_ARGUMENT, "NULL argument");
    if (root < 0 || root >= c->ngpu || count < 0) SK_FAIL(SK_ERR_ARGUMENT, "bad root or count");
    for (int g = 0; g < c->ngpu; g++) if (!theta_dev[g]) SK_FAIL(SK_ERR_ARGUMENT, "NULL device buffer");
    const NcclApi &n = nccl();
    int prev = 0;
    cudaGetDevice(&prev);
    NK(n.GroupStart());
    for (int g = 0; g < c->ngpu; g++) {
        cudaSetDevice(c->devices[(size_t)g]);
        ncclResult_t r = n.Broadcast(theta_dev[g], theta_dev[g], (size_t)count, ncclFloat64_, root, c->comms[(size_t)g], c->streams[(size_t)g]);
        if (r != ncclSuccess_) { n.GroupEnd(); cudaSetDevice(prev); NK(r); }
    }
    NK(n.GroupEnd());
    for (int g = 0; g < c->ngpu; g++) { cudaSetDevice(c->devices[(size_t)g]); cudaStreamSynchronize(c->streams[(size_t)g]); }
    cudaSetDevice(prev);
    skk::count_launch(c->ngpu);
    return SK_OK;
}

// linear_combination over device-resident shards.  Slices are contiguous, ceil(n/G) points each (the rule of
// sk_linear_combination_sharded): device g holds points [g·per, min(n, (g+1)·per)) in x_dev[g] and its own copy of θ
// in theta_dev[g].  gather == 0: out_dev[g] receives that slice's results ([slice][E]).  gather != 0 (C2): every
// out_dev[g] is a full buffer of per·G rows; device g writes its slice at row g·per and an in-place ncclAllGather
// completes all of them (rows >= n are padding).  ms_compute / ms_gather (optional, [G]) return the device-timed
// duration of the two phases on every device.
int sk_linear_combination_sharded_device(sk_comm *c, sk_plan *const *plans, const double *const *theta_dev, int64_t theta_len,
                                         int nvec, const double *const *x_dev, int64_t n, double *const *out_dev, int gather,
                                         uint32_t flags, float *ms_compute, float *ms_gather) {
    if (!c || !plans || !theta_dev || !x_dev || !out_dev) SK_FAIL(SK_ERR_ARGUMENT, "NULL argument");
    if (n < 0 || nvec < 1) SK_FAIL(SK_ERR_ARGUMENT, "bad n or nvec");
    const int G = c->ngpu;
    sk_plan_info i0{};
    for (int g = 0; g < G; g++) {
        if (!plans[g]) SK_FAIL(SK_ERR_ARGUMENT, "plan is NULL");
        sk_plan_info ig{};
        int rc = sk_plan_get_info(plans[g], &ig);
        if (rc != SK_OK) return rc;
        if (ig.device != c->devices[(size_t)g]) SK_FAIL(SK_ERR_ARGUMENT, "plans[g] must live on the communicator's device g");
        if (g == 0) i0 = ig;
        else if (ig.K != i0.K || ig.E != i0.E || ig.d != i0.d) SK_FAIL(SK_ERR_ARGUMENT, "plans must describe the same basis");
    }
    const int Eout = nvec > 1 ? nvec : i0.E;
    const int64_t per = (n + G - 1) / G;
    std::vector<int> rcs((size_t)G, SK_OK);
    std::vector<std::string> msgs((size_t)G);
    std::vector<std::thread> workers;
    if (gather == 2 && G > 1 && c->peer_ok) {
        // C2 overlapped with the compute: every device cuts its slice into chunks; as soon as the kernel of a chunk has
        // finished, the copy engines push that chunk's results into the full buffers of the other devices (one stream per
        // destination, peer-to-peer over NVLink / NVSwitch) while the SMs evaluate the next chunk.  No NCCL, no second pass
        // over the result; what remains after the last kernel is the transfer of one chunk.
        const int64_t target = 1 << 20;                                   // points per chunk: ≈ 0.7 ms of the headline kernel
        const int C = (int)std::max<int64_t>(1, std::min<int64_t>(16, per / target));
        std::vector<std::vector<cudaEvent_t>> evc((size_t)G, std::vector<cudaEvent_t>((size_t)C, nullptr));
        std::vector<std::vector<cudaEvent_t>> evd((size_t)G, std::vector<cudaEvent_t>((size_t)G, nullptr));
        for (int g = 0; g < G; g++) {
            workers.emplace_back([&, g]() {
                const int64_t lo = std::min<int64_t>(n, per * g), hi = std::min<int64_t>(n, lo + per);
                const int64_t cs = (hi - lo + C - 1) / std::max(C, 1);
                cudaSetDevice(c->devices[(size_t)g]);
                cudaStream_t sg = c->streams[(size_t)g];
                cudaEventRecord(c->ev0[(size_t)g], sg);
                // FUSED: where the plan's kernel has a peer-storing instantiation (the six-dimensional leaf kernel: cfg 4),
                // the threads that compute a point's results store them into every device's buffer themselves — one launch,
                // no copy at all; NVLink carries the stores while the FP64 pipe works on the next points
                bool fused = false;
                if (hi > lo && nvec == 1 && !(flags & SK_MODE_EXACT) && G <= 8 && skk::smolyak_leaf_serves_peers(*plans[g])) {
                    double *peers[7];
                    int np = 0;
                    for (int h = 0; h < G; h++) if (h != g) peers[np++] = out_dev[h] + (size_t)lo * Eout;
                    skk::smolyak_leaf_set_peers(np, peers);
                    rcs[(size_t)g] = sk_linear_combination(plans[g], theta_dev[g], theta_len, nvec, x_dev[g], hi - lo, out_dev[g] + (size_t)lo * Eout,
                                                           (flags & ~SK_MEM_HOST) | SK_MEM_DEVICE, sg);
                    fused = skk::smolyak_leaf_peers_used();
                    skk::smolyak_leaf_set_peers(0, nullptr);
                    if (rcs[(size_t)g] != SK_OK) msgs[(size_t)g] = sk_last_error();
                }
                for (int j = 0; j < C && rcs[(size_t)g] == SK_OK && !fused; j++) {
                    const int64_t clo = std::min<int64_t>(hi, lo + cs * j), chi = std::min<int64_t>(hi, clo + cs);
                    if (chi <= clo) break;
                    double *dst = out_dev[g] + (size_t)clo * Eout;
                    rcs[(size_t)g] = sk_linear_combination(plans[g], theta_dev[g], theta_len, nvec, x_dev[g] + (size_t)(clo - lo) * i0.d, chi - clo, dst,
                                                           (flags & ~SK_MEM_HOST) | SK_MEM_DEVICE, sg);
                    if (rcs[(size_t)g] != SK_OK) { msgs[(size_t)g] = sk_last_error(); break; }
                    cudaEventCreateWithFlags(&evc[(size_t)g][(size_t)j], cudaEventDisableTiming);
                    cudaEventRecord(evc[(size_t)g][(size_t)j], sg);
                    for (int h = 0; h < G; h++) {
                        if (h == g) continue;
                        cudaStream_t sc = c->cstreams[(size_t)g][(size_t)h];
                        cudaStreamWaitEvent(sc, evc[(size_t)g][(size_t)j], 0);
                        cudaMemcpyPeerAsync(out_dev[h] + (size_t)clo * Eout, c->devices[(size_t)h], dst, c->devices[(size_t)g],
                                            (size_t)(chi - clo) * Eout * sizeof(double), sc);
                    }
                }
                cudaEventRecord(c->ev1[(size_t)g], sg);                   // end of this device's kernels
                for (int h = 0; h < G; h++) {                             // the call's stream finishes after its copies
                    if (h == g) continue;
                    cudaEventCreateWithFlags(&evd[(size_t)g][(size_t)h], cudaEventDisableTiming);
                    cudaEventRecord(evd[(size_t)g][(size_t)h], c->cstreams[(size_t)g][(size_t)h]);
                    cudaStreamWaitEvent(sg, evd[(size_t)g][(size_t)h], 0);
                }
            });
        }
        for (auto &w : workers) w.join();
        int prev2 = 0;
        cudaGetDevice(&prev2);
        std::vector<cudaEvent_t> tail((size_t)G, nullptr);
        for (int g = 0; g < G; g++) { cudaSetDevice(c->devices[(size_t)g]); cudaEventCreate(&tail[(size_t)g]); cudaEventRecord(tail[(size_t)g], c->streams[(size_t)g]); }
        cudaError_t e2 = cudaSuccess;
        for (int g = 0; g < G; g++) {
            cudaSetDevice(c->devices[(size_t)g]);
            cudaError_t eg = cudaStreamSynchronize(c->streams[(size_t)g]);
            if (eg != cudaSuccess && e2 == cudaSuccess) e2 = eg;
        }
        for (int g = 0; g < G; g++) {
            cudaSetDevice(c->devices[(size_t)g]);
            if (e2 == cudaSuccess && ms_compute) cudaEventElapsedTime(&ms_compute[g], c->ev0[(size_t)g], c->ev1[(size_t)g]);
            if (ms_gather) ms_gather[g] = 0.f;
            if (e2 == cudaSuccess && ms_gather) cudaEventElapsedTime(&ms_gather[g], c->ev1[(size_t)g], tail[(size_t)g]);      // what the copies add after the last kernel
            cudaEventDestroy(tail[(size_t)g]);
            for (cudaEvent_t ev : evc[(size_t)g]) if (ev) cudaEventDestroy(ev);
            for (cudaEvent_t ev : evd[(size_t)g]) if (ev) cudaEventDestroy(ev);
        }
        cudaSetDevice(prev2);
        for (int g = 0; g < G; g++) if (rcs[(size_t)g] != SK_OK) SK_FAIL(rcs[(size_t)g], "device " + std::to_string(c->devices[(size_t)g]) + ": " + msgs[(size_t)g]);
        CKM(e2);
        skk::count_launch(G * (G - 1) * C);
        return SK_OK;
    }
    // phase 1: every device evaluates its slice on its own stream (one host thread per device: launches overlap)
    for (int g = 0; g < G; g++) {
        workers.emplace_back([&, g]() {
            const int64_t lo = std::min<int64_t>(n, per * g), hi = std::min<int64_t>(n, lo + per);
            cudaSetDevice(c->devices[(size_t)g]);
            cudaEventRecord(c->ev0[(size_t)g], c->streams[(size_t)g]);
            if (hi > lo) {
                double *dst = gather ? out_dev[g] + (size_t)lo * Eout : out_dev[g];
                rcs[(size_t)g] = sk_linear_combination(plans[g], theta_dev[g], theta_len, nvec, x_dev[g], hi - lo, dst,
                                                       (flags & ~SK_MEM_HOST) | SK_MEM_DEVICE, c->streams[(size_t)g]);
                if (rcs[(size_t)g] != SK_OK) msgs[(size_t)g] = sk_last_error();
            }
            cudaEventRecord(c->ev1[(size_t)g], c->streams[(size_t)g]);
        });
    }
    for (auto &w : workers) w.join();
    for (int g = 0; g < G; g++) if (rcs[(size_t)g] != SK_OK) SK_FAIL(rcs[(size_t)g], "device " + std::to_string(c->devices[(size_t)g]) + ": " + msgs[(size_t)g]);
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<cudaEvent_t> g0, g1;
    if (gather && G > 1) {
        // phase 2 (C2): in-place all-gather of the slices, stream-ordered after each device's kernel
        const NcclApi &nc = nccl();
        g0.assign((size_t)G, nullptr); g1.assign((size_t)G, nullptr);
        for (int g = 0; g < G; g++) { cudaSetDevice(c->devices[(size_t)g]); cudaEventCreate(&g0[(size_t)g]); cudaEventCreate(&g1[(size_t)g]); cudaEventRecord(g0[(size_t)g], c->streams[(size_t)g]); }
        ncclResult_t r = nc.GroupStart();
        for (int g = 0; g < G && r == ncclSuccess_; g++) {
            cudaSetDevice(c->devices[(size_t)g]);
            r = nc.AllGather(out_dev[g] + (size_t)per * g * Eout, out_dev[g], (size_t)per * Eout, ncclFloat64_, c->comms[(size_t)g], c->streams[(size_t)g]);
        }
        ncclResult_t r2 = nc.GroupEnd();
        for (int g = 0; g < G; g++) { cudaSetDevice(c->devices[(size_t)g]); cudaEventRecord(g1[(size_t)g], c->streams[(size_t)g]); }
        if (r != ncclSuccess_ || r2 != ncclSuccess_) {
            for (int g = 0; g < G; g++) { cudaEventDestroy(g0[(size_t)g]); cudaEventDestroy(g1[(size_t)g]); }
            cudaSetDevice(prev);
            NK(r != ncclSuccess_ ? r : r2);
        }
        skk::count_launch(G);
    }
    cudaError_t e = cudaSuccess;
    for (int g = 0; g < G; g++) {
        cudaSetDevice(c->devices[(size_t)g]);
        cudaError_t eg = cudaStreamSynchronize(c->streams[(size_t)g]);
        if (eg != cudaSuccess && e == cudaSuccess) e = eg;
        if (eg == cudaSuccess && ms_compute) cudaEventElapsedTime(&ms_compute[g], c->ev0[(size_t)g], c->ev1[(size_t)g]);
        if (ms_gather) ms_gather[g] = 0.f;
        if (eg == cudaSuccess && ms_gather && !g0.empty()) cudaEventElapsedTime(&ms_gather[g], g0[(size_t)g], g1[(size_t)g]);
        if (!g0.empty()) { cudaEventDestroy(g0[(size_t)g]); cudaEventDestroy(g1[(size_t)g]); }
    }
    cudaSetDevice(prev);
    CKM(e);
    return SK_OK;
}

int sk_linear_combination_peers(const sk_plan *plan, const double *theta, int64_t theta_len, const double *x, int64_t n,
                                double *out, double *const *peer_out, int n_peer, uint32_t flags, void *stream, int *fused) {
    if (fused) *fused = 0;
    if (!plan || (n_peer > 0 && !peer_out)) SK_FAIL(SK_ERR_ARGUMENT, "NULL argument");
    if (n_peer < 0 || n_peer > 7) SK_FAIL(SK_ERR_ARGUMENT, "between 0 and 7 peer buffers");
    if (!(flags & SK_MEM_DEVICE)) SK_FAIL(SK_ERR_ARGUMENT, "sk_linear_combination_peers takes device pointers (SK_MEM_DEVICE)");
    for (int g = 0; g < n_peer; g++) if (!peer_out[g]) SK_FAIL(SK_ERR_ARGUMENT, "NULL peer buffer");
    const bool can = n_peer > 0 && n > 0 && !(flags & SK_MODE_EXACT) && skk::smolyak_leaf_serves_peers(*plan);
    if (can) skk::smolyak_leaf_set_peers(n_peer, peer_out);
    const int rc = sk_linear_combination(plan, theta, theta_len, 1, x, n, out, flags, stream);
    if (can) {
        if (fused) *fused = skk::smolyak_leaf_peers_used() ? 1 : 0;
        skk::smolyak_leaf_set_peers(0, nullptr);
    }
    return rc;
}

}  // extern "C"
