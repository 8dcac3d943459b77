// api.cu -- the C ABI of libb200kzg.so (include/b200_kzg.h): handles, host-side level-1
// operations of package bls, and the orchestration of the device pipelines (Fr NTT, G1 FFT,
// LinCombG1, FK20 single / multi).  Every compute entry point runs on the GPU; there is no CPU
// fallback (a missing device is an error).
#include <cuda_runtime.h>
#include <atomic>
#include <stdlib.h>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>
#include "../../include/b200_kzg.h"
#include "hostutil.cuh"
#include "pairing.h"
#include "kernels.h"

using namespace b200;

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_cuda_err;
static thread_local int g_device = 0;

#define CK(expr)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            g_cuda_err = std::string(#expr) + ": " + cudaGetErrorString(e__);             \
            return e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? B200_ERR_NO_DEVICE : B200_ERR_CUDA; \
        }                                                                                 \
    } while (0)
#define CKS(expr)                       \
    do {                                \
        int s__ = (expr);               \
        if (s__ != B200_OK) return s__; \
    } while (0)

extern "C" const char* b200_strerror(int status) {
    switch (status) {
        case B200_OK: return "ok";
        case B200_ERR_TOO_LARGE: return "more values than roots of unity";
        case B200_ERR_NOT_POW2: return "not a power of two";
        case B200_ERR_LEN_MISMATCH: return "length mismatch";
        case B200_ERR_BAD_INPUT: return "bad input";
        case B200_ERR_CUDA: return "CUDA error";
        case B200_ERR_NO_DEVICE: return "no CUDA device";
        case B200_ERR_TOO_SMALL: return "too small";
        case B200_ERR_RECOVERY: return "recovered data does not match the known samples";
        case B200_ERR_ZERO_EVAL: return "bad zero eval";
    }
    return "unknown status";
}
extern "C" const char* b200_last_cuda_error(void) { return g_cuda_err.c_str(); }
extern "C" int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int b200_set_device(int device) {
    int n = b200_device_count();
    if (n == 0) return B200_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return B200_ERR_BAD_INPUT;
    g_device = device;
    return B200_OK;
}

// sticky-error-free check after a batch of launches
static int check_launches() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_cuda_err = std::string("kernel launch: ") + cudaGetErrorString(e); return B200_ERR_CUDA; }
    return B200_OK;
}

// Work buffers come from the device's stream-ordered pool.  Its default release threshold (0)
// hands every freed block back to the driver at the next synchronisation, so a host-buffer call
// (which ends in cudaStreamSynchronize) would re-map hundreds of MB on every call; keep them.
static void keep_pool_memory() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    done[dev] = true;
}

// Host-buffer entry points run on a pooled non-blocking stream leased for the duration of the call (one pool per
// device): callers on different threads (goroutines migrate across OS threads) overlap instead of serialising on
// the legacy default stream.  Streams are never destroyed; a lease taken after cudaSetDevice belongs to that device.
// Calls in flight (host-buffer entry points).  A call that finds the device to itself may spend lanes on latency (a quad of
// lanes per butterfly in small G1 transforms, kernels_g1.cu); with several callers in flight the SMs are busy anyway and the
// one-lane-per-butterfly kernels give more throughput.  Decided once per call, on the calling thread.
static std::atomic<int> g_active_calls{0};
static std::atomic<int> g_latency_mode{1};      // b200_set_latency_mode: 1 = automatic (default), 0 = never
static const int kQuadMaxCallers = 2;
extern "C" int b200_set_latency_mode(int mode) {
    if (mode != 0 && mode != 1) return B200_ERR_BAD_INPUT;
    g_latency_mode = mode;
    g1_set_quad_enabled(mode == 1);
    return B200_OK;
}

struct StreamLease {
    cudaStream_t st = nullptr;
    int dev = -1;
    bool counted = false;
    void count() { counted = true; g1_set_quad_allowed(++g_active_calls <= kQuadMaxCallers && g_latency_mode.load() == 1); }
    static std::mutex& mu() { static std::mutex m; return m; }
    static std::vector<cudaStream_t>& pool(int d) { static std::vector<cudaStream_t> p[64]; return p[d]; }
    int acquire(bool counted_call = true) {      // counted_call = false: an extra stream of a call that already holds one
        CK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64) return B200_ERR_BAD_INPUT;
        {
            std::lock_guard<std::mutex> lk(mu());
            auto& p = pool(dev);
            if (!p.empty()) { st = p.back(); p.pop_back(); if (counted_call) count(); return B200_OK; }
        }
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (counted_call) count();
        return B200_OK;
    }
    ~StreamLease() {
        if (counted) --g_active_calls;
        if (!st) return;
        std::lock_guard<std::mutex> lk(mu());
        pool(dev).push_back(st);
    }
};

// one point at infinity per device (source of strided clears); never freed
static int dev_infinity(G1J** out) {
    static std::mutex mu;
    static G1J* inf[64] = {};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return B200_ERR_BAD_INPUT;
    std::lock_guard<std::mutex> lk(mu);
    if (!inf[dev]) { CK(cudaMalloc(&inf[dev], sizeof(G1J))); CK(cudaMemset(inf[dev], 0, sizeof(G1J))); CK(cudaDeviceSynchronize()); }
    *out = inf[dev];
    return B200_OK;
}

// RAII stream-ordered device buffer
struct DevBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    int alloc(size_t bytes, cudaStream_t s) {
        st = s;
        if (bytes == 0) bytes = 16;
        keep_pool_memory();
        CK(cudaMallocAsync(&p, bytes, s));
        return B200_OK;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static inline bool is_pow2(uint64_t v) { return v && !(v & (v - 1)); }   // bls/globals.go:72
static inline unsigned log2u(uint64_t v) { unsigned l = 0; while (((uint64_t)1 << l) < v) l++; return l; }
static inline uint64_t next_pow2(uint64_t v) { if (v == 0) return 1; return (uint64_t)1 << log2u(v); }   // fft.go:11-16

// ------------------------------------------------------------------------------ level 1 (host)
extern "C" void b200_fr_add(uint64_t* dst, const uint64_t* a, const uint64_t* b) { fr_store_canon(dst, fe_add(fr_load_canon(a), fr_load_canon(b))); }
extern "C" void b200_fr_sub(uint64_t* dst, const uint64_t* a, const uint64_t* b) { fr_store_canon(dst, fe_sub(fr_load_canon(a), fr_load_canon(b))); }
extern "C" void b200_fr_mul(uint64_t* dst, const uint64_t* a, const uint64_t* b) {
    // canonical x canonical: (a R)(b) / R = a b
    fr_store_canon(dst, fe_mul(fe_to_mont(fr_load_canon(a)), fr_load_canon(b)));
}
extern "C" void b200_fr_inv(uint64_t* dst, const uint64_t* a) { fr_to_abi_from_mont(dst, fe_inv(fr_from_abi_mont(a))); }
extern "C" void b200_fr_div(uint64_t* dst, const uint64_t* a, const uint64_t* b) {   // bls/bignum_kilic.go:103-107
    Fr bi = fe_inv(fr_from_abi_mont(b));
    fr_store_canon(dst, fe_mul(bi, fr_load_canon(a)));
}
extern "C" void b200_fr_batch_inv(uint64_t* vals, size_t n) {   // Montgomery's trick; zeros stay zero
    if (!n) return;
    std::vector<Fr> v(n), pre(n);
    Fr acc = Fr::one();
    for (size_t i = 0; i < n; i++) {
        v[i] = fr_from_abi_mont(vals + 4 * i);
        pre[i] = acc;
        if (!v[i].is_zero()) acc = fe_mul(acc, v[i]);
    }
    acc = fe_inv(acc);
    for (size_t i = n; i-- > 0;) {
        if (v[i].is_zero()) { fr_store_canon(vals + 4 * i, Fr::zero()); continue; }
        Fr inv = fe_mul(acc, pre[i]);
        acc = fe_mul(acc, v[i]);
        fr_to_abi_from_mont(vals + 4 * i, inv);
    }
}
extern "C" int b200_fr_valid(const uint8_t* le32) {
    uint64_t l[4];
    memcpy(l, le32, 32);
    return fr_canon_valid(l) ? 1 : 0;
}
extern "C" void b200_fr_root_of_unity(unsigned scale, uint64_t* out) {
    if (scale > 31) { memset(out, 0, 32); return; }
    fr_store_canon(out, fr_scale2_root_canon(scale));
}
extern "C" void b200_g1_generator(uint64_t* out) { g1_to_abi(out, g1_generator()); }
extern "C" void b200_g1_add(uint64_t* dst, const uint64_t* a, const uint64_t* b) { g1_to_abi(dst, g1_add(g1_from_abi(a), g1_from_abi(b))); }
extern "C" void b200_g1_sub(uint64_t* dst, const uint64_t* a, const uint64_t* b) { g1_to_abi(dst, g1_sub(g1_from_abi(a), g1_from_abi(b))); }
extern "C" void b200_g1_neg(uint64_t* dst) { g1_to_abi(dst, g1_neg(g1_from_abi(dst))); }
extern "C" void b200_g1_mul(uint64_t* dst, const uint64_t* a, const uint64_t* k) {
    Fr s = fr_load_canon(k);
    g1_to_abi(dst, g1_mul_simple(g1_from_abi(a), s.l));
}
extern "C" int b200_g1_equal(const uint64_t* a, const uint64_t* b) { return g1_equal(g1_from_abi(a), g1_from_abi(b)) ? 1 : 0; }
extern "C" void b200_g1_to_compressed(uint8_t* out, const uint64_t* p) { g1_compress(out, g1_from_abi(p)); }
extern "C" int b200_g1_from_compressed(uint64_t* out, const uint8_t* in) {
    G1J p;
    int rc = g1_decompress(p, in);
    if (rc) return B200_ERR_BAD_INPUT;
    g1_to_abi(out, p);
    return B200_OK;
}
extern "C" void b200_g1_to_compressed_many(uint8_t* out, const uint64_t* pts, size_t n) {
    for (size_t i = 0; i < n; i++) b200_g1_to_compressed(out + 48 * i, pts + 18 * i);
}
extern "C" int b200_g1_from_compressed_many(uint64_t* out, const uint8_t* in, size_t n) {
    for (size_t i = 0; i < n; i++) CKS(b200_g1_from_compressed(out + 18 * i, in + 48 * i));
    return B200_OK;
}

// ---- G2 and the pairing (pairing.h): host code, verification side only -------------------------------------------
extern "C" void b200_g2_generator(uint64_t* out) { g2_to_abi(out, g2_generator()); }
extern "C" void b200_g2_add(uint64_t* dst, const uint64_t* a, const uint64_t* b) { g2_to_abi(dst, g2_add(g2_from_abi(a), g2_from_abi(b))); }
extern "C" void b200_g2_sub(uint64_t* dst, const uint64_t* a, const uint64_t* b) { g2_to_abi(dst, g2_sub(g2_from_abi(a), g2_from_abi(b))); }
extern "C" void b200_g2_neg(uint64_t* dst) { g2_to_abi(dst, g2_neg(g2_from_abi(dst))); }
extern "C" void b200_g2_mul(uint64_t* dst, const uint64_t* a, const uint64_t* k) { g2_to_abi(dst, g2_mul(g2_from_abi(a), fr_load_canon(k))); }
extern "C" int b200_g2_equal(const uint64_t* a, const uint64_t* b) { return g2_equal(g2_from_abi(a), g2_from_abi(b)) ? 1 : 0; }
extern "C" void b200_g2_to_compressed(uint8_t* out, const uint64_t* p) { g2_compress(out, g2_from_abi(p)); }
extern "C" int b200_g2_from_compressed(uint64_t* out, const uint8_t* in) {
    G2J p;
    if (g2_decompress(p, in)) return B200_ERR_BAD_INPUT;
    g2_to_abi(out, p);
    return B200_OK;
}
// affine coordinates (canonical, x then y; zeros for infinity) -- StrG1 / StrG2 print them (bls/bls_kilic.go:55-61,96-102)
extern "C" void b200_g1_to_affine(const uint64_t* p, uint64_t* xy) {
    memset(xy, 0, 96);
    const G1J a = g1_from_abi(p);
    if (a.is_inf()) return;
    Fp x, y;
    g1_affine_canon(a, x, y);
    memcpy(xy, x.l, 48); memcpy(xy + 6, y.l, 48);
}
extern "C" void b200_g2_to_affine(const uint64_t* p, uint64_t* xy) {
    memset(xy, 0, 192);
    const G2J a = g2_from_abi(p);
    if (a.is_inf()) return;
    Fp2 x, y;
    g2_affine(a, x, y);
    const Fp c[4] = {fe_from_mont(x.c0), fe_from_mont(x.c1), fe_from_mont(y.c0), fe_from_mont(y.c1)};
    for (int i = 0; i < 4; i++) memcpy(xy + 6 * i, c[i].l, 48);
}
// setup.go:9-26 GenerateTestingSetup, G2 half: out[i] = secret^i * GenG2 (host: ~n scalar multiplications)
extern "C" int b200_generate_testing_setup_g2(const uint64_t* secret, size_t n, uint64_t* out) {
    const Fr s = fr_from_abi_mont(secret);
    Fr pw = Fr::one();
    const G2J g = g2_generator();
    for (size_t i = 0; i < n; i++) {
        g2_to_abi(out + 36 * i, g2_mul(g, fe_from_mont(pw)));
        pw = fe_mul(pw, s);
    }
    return B200_OK;
}
static bool abi_coords_canonical(const uint64_t* p, int coords) {
    for (int i = 0; i < coords; i++) if (!fp_abi_canonical(p + 6 * i)) return false;
    return true;
}
// bls/bls_kilic.go:152-158 PairingsVerify: e(a1, a2) == e(b1, b2).  Points off their curve -> B200_ERR_BAD_INPUT
// (kilic's engine assumes valid points; an answer for invalid ones would be meaningless).
extern "C" int b200_pairings_verify(const uint64_t* a1, const uint64_t* a2, const uint64_t* b1, const uint64_t* b2, int* ok) {
    *ok = 0;
    if (!abi_coords_canonical(a1, 3) || !abi_coords_canonical(b1, 3) || !abi_coords_canonical(a2, 6) || !abi_coords_canonical(b2, 6))
        return B200_ERR_BAD_INPUT;
    const G1J p1 = g1_from_abi(a1), p2 = g1_from_abi(b1);
    const G2J q1 = g2_from_abi(a2), q2 = g2_from_abi(b2);
    if (!g1_on_curve(p1) || !g1_on_curve(p2) || !g2_on_curve(q1) || !g2_on_curve(q2)) return B200_ERR_BAD_INPUT;
    *ok = pairings_verify(p1, q1, p2, q2) ? 1 : 0;
    return B200_OK;
}
// e(p, q) as the 12 canonical Fp coefficients (a_0, b_0, .., a_5, b_5) of sum (a_i + b_i u) w^i, w^6 = 1 + u
extern "C" int b200_pairing(const uint64_t* p, const uint64_t* q, uint64_t* out) {
    if (!abi_coords_canonical(p, 3) || !abi_coords_canonical(q, 6)) return B200_ERR_BAD_INPUT;
    const G1J pp = g1_from_abi(p);
    const G2J qq = g2_from_abi(q);
    if (!g1_on_curve(pp) || !g2_on_curve(qq)) return B200_ERR_BAD_INPUT;
    const Fp12 e = pairing(pp, qq);
    for (int i = 0; i < 6; i++) {
        Fp a = fe_from_mont(e.c[i].c0), b = fe_from_mont(e.c[i].c1);
        memcpy(out + 12 * i, a.l, 48); memcpy(out + 12 * i + 6, b.l, 48);
    }
    return B200_OK;
}

// ------------------------------------------------------------------------------ FFTSettings
struct b200_fs {
    int device = 0;
    unsigned max_scale = 0;
    uint64_t max_width = 0;
    std::vector<Fr> h_expanded;   // Montgomery, max_width + 1
    FrDomain dom;
    std::mutex mu;
    // twiddle programs for the G1 FFT: [inverse][mode], max_width / 2 entries each, built lazily
    ScalarProgram* progs[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    Fr* d_shift[2] = {nullptr, nullptr};   // 5^-i and 5^i, i < max_width (recovery), built lazily
};

static Fr fr_inv_of_u64(uint64_t v) { return fe_inv(fr_from_u64(v)); }   // Montgomery in, Montgomery out

extern "C" int b200_fft_settings_new(uint8_t max_scale, b200_fs** out) {
    *out = nullptr;
    if (max_scale > 31) return B200_ERR_TOO_LARGE;
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    CK(cudaSetDevice(g_device));
    b200_fs* fs = new (std::nothrow) b200_fs();
    if (!fs) return B200_ERR_CUDA;
    fs->device = g_device;
    fs->max_scale = max_scale;
    fs->max_width = (uint64_t)1 << max_scale;
    const uint64_t W = fs->max_width;
    // fft.go:21-32 expandRootOfUnity: 1, w, w^2, ..., w^W = 1
    Fr w = fe_to_mont(fr_scale2_root_canon(max_scale));
    fs->h_expanded.resize(W + 1);
    fs->h_expanded[0] = Fr::one();
    for (uint64_t i = 1; i <= W; i++) fs->h_expanded[i] = fe_mul(fs->h_expanded[i - 1], w);
    std::vector<Fr> rev(W + 1), twf(W ? W : 1), twi(W ? W : 1);
    for (uint64_t i = 0; i <= W; i++) rev[i] = fs->h_expanded[W - i];
    twf[0] = twi[0] = Fr::one();
    for (uint64_t n = 2; n <= W; n <<= 1) {
        uint64_t stride = W / n;
        for (uint64_t j = 0; j < n / 2; j++) {
            twf[n / 2 + j] = fs->h_expanded[j * stride];
            twi[n / 2 + j] = rev[j * stride];
        }
    }
    FrDomain& d = fs->dom;
    d.max_scale = max_scale; d.max_width = W;
    auto up = [&](Fr** dst, const std::vector<Fr>& src) -> int {
        CK(cudaMalloc(dst, src.size() * sizeof(Fr)));
        CK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(Fr), cudaMemcpyHostToDevice));
        return B200_OK;
    };
    // cudaMemcpy from pageable memory may return before the DMA has landed; the calls that use these tables run on
    // non-blocking streams with no implicit ordering against it, hence the device-wide synchronisation below
    int rc = up(&d.expanded, fs->h_expanded);
    if (!rc) rc = up(&d.reverse, rev);
    if (!rc) rc = up(&d.tw_fwd, twf);
    if (!rc) rc = up(&d.tw_inv, twi);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) { g_cuda_err = "sync after the domain tables"; rc = B200_ERR_CUDA; }
    if (rc) { b200_fft_settings_free(fs); return rc; }
    *out = fs;
    return B200_OK;
}
extern "C" void b200_fft_settings_free(b200_fs* fs) {
    if (!fs) return;
    cudaSetDevice(fs->device);
    cudaFree(fs->dom.expanded); cudaFree(fs->dom.reverse); cudaFree(fs->dom.tw_fwd); cudaFree(fs->dom.tw_inv);
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) cudaFree(fs->progs[a][b]);
    cudaFree(fs->d_shift[0]); cudaFree(fs->d_shift[1]);
    delete fs;
}
extern "C" uint64_t b200_fs_max_width(const b200_fs* fs) { return fs->max_width; }
extern "C" int b200_fs_roots(const b200_fs* fs, int reverse, uint64_t* out) {
    const uint64_t W = fs->max_width;
    for (uint64_t i = 0; i <= W; i++) fr_to_abi_from_mont(out + 4 * i, fs->h_expanded[reverse ? W - i : i]);
    return B200_OK;
}

// twiddle programs w^(+-j), j < max_width / 2, in the given recoding mode
static int fs_programs(b200_fs* fs, int inverse, int mode, const ScalarProgram** out) {
    std::lock_guard<std::mutex> lk(fs->mu);
    if (!fs->progs[inverse][mode]) {
        const uint64_t W = fs->max_width, half = W / 2 ? W / 2 : 1;
        std::vector<ScalarProgram> h(half);
        for (uint64_t j = 0; j < half; j++) {
            Fr k = fe_from_mont(fs->h_expanded[inverse ? (W - j) % (W ? W : 1) : j]);
            if (W == 0) k = fe_from_mont(Fr::one());
            make_scalar_program(&h[j], k, mode);
        }
        ScalarProgram* d = nullptr;
        CK(cudaMalloc(&d, half * sizeof(ScalarProgram)));
        CK(cudaMemcpy(d, h.data(), half * sizeof(ScalarProgram), cudaMemcpyHostToDevice));
        CK(cudaDeviceSynchronize());   // the copy must have landed before a non-blocking stream reads the table
        fs->progs[inverse][mode] = d;
    }
    *out = fs->progs[inverse][mode];
    return B200_OK;
}
// lanes of a warp share the twiddle when the batch fills whole warps -> sparse (wNAF) recoding
// matches g1_lanes_for_batch (kernels_g1.cu): whole warps per butterfly from 16 blobs on
static inline int program_mode_for_batch(size_t batch) { return batch >= 16 ? 1 : 0; }

// ------------------------------------------------------------------------------ Fr FFT
static int dev_fr_fft(b200_fs* fs, const Fr* d_in, Fr* d_out, unsigned logn, size_t batch, bool inverse, cudaStream_t st) {
    if (logn > 24) return B200_ERR_TOO_LARGE;   // two passes of at most 2^12 points each (kernels_fr.cu: launch_fr_ntt)
    DevBuf tmp;
    if (logn > 12) CKS(tmp.alloc(((size_t)batch << logn) * sizeof(Fr), st));
    Fr scale;
    if (inverse) scale = fr_inv_of_u64((uint64_t)1 << logn);
    launch_fr_ntt(fs->dom, d_in, d_out, tmp.as<Fr>(), logn, batch, inverse, inverse ? &scale : nullptr, st);
    return check_launches();
}

extern "C" int b200_fft_fr_batch(b200_fs* fs, const uint64_t* vals, size_t n, size_t batch, int inverse, uint64_t* out) {
    if (n > fs->max_width) return B200_ERR_TOO_LARGE;              // fft_fr.go:57-59
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fs->device));
    const uint64_t np = next_pow2(n);                              // fft_fr.go:60
    const unsigned logn = log2u(np);
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf raw, buf;
    CKS(raw.alloc(batch * np * 32, st));
    CKS(buf.alloc(batch * np * sizeof(Fr), st));
    if (np != n) CK(cudaMemsetAsync(raw.p, 0, batch * np * 32, st));   // zero padding fft_fr.go:66-68
    if (n) CK(cudaMemcpy2DAsync(raw.p, np * 32, vals, n * 32, n * 32, batch, cudaMemcpyHostToDevice, st));
    launch_fr_to_mont(raw.as<uint64_t>(), buf.as<Fr>(), batch * np, st);
    CKS(dev_fr_fft(fs, buf.as<Fr>(), buf.as<Fr>(), logn, batch, inverse != 0, st));
    launch_fr_from_mont(buf.as<Fr>(), raw.as<uint64_t>(), batch * np, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out, raw.p, batch * np * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
extern "C" int b200_fft_fr(b200_fs* fs, const uint64_t* vals, size_t n, int inverse, uint64_t* out) {
    return b200_fft_fr_batch(fs, vals, n, 1, inverse, out);
}
// fft_fr.go:76-105 InplaceFFT: no padding -- a length that is not a power of two is an error (:81-83); n == 0
// passes both checks and then divides by n (:89,100), a run-time panic.  vals and out may be the same buffer.
extern "C" int b200_inplace_fft_fr(b200_fs* fs, const uint64_t* vals, size_t n, int inverse, uint64_t* out) {
    if (n > fs->max_width) return B200_ERR_TOO_LARGE;              // fft_fr.go:78-80
    if (n == 0) return B200_ERR_BAD_INPUT;
    if (!is_pow2(n)) return B200_ERR_NOT_POW2;                     // fft_fr.go:81-83
    return b200_fft_fr_batch(fs, vals, n, 1, inverse, out);
}

// ------------------------------------------------------------------------------ DAS extension
extern "C" int b200_das_fft_extension_batch(b200_fs* fs, uint64_t* vals, size_t n, size_t batch) {
    if (n * 2 > fs->max_width) return B200_ERR_TOO_SMALL;   // das_extension.go:72-74
    if (n < 2 || !is_pow2(n)) return B200_ERR_BAD_INPUT;    // das_extension.go:22-24 "bad usage"
    if (batch == 0) return B200_OK;
    if (batch > 65535) return B200_ERR_TOO_LARGE;           // the batch rides on grid.y
    CK(cudaSetDevice(fs->device));
    // (cutting the batch into chunks on several streams to overlap the copies was measured SLOWER: 55 k against 69 k
    // polynomials/s at n = 8192, batch 64 -- the per-chunk allocations and launches cost more than the overlap returns)
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf raw, buf;
    CKS(raw.alloc(batch * n * 32, st)); CKS(buf.alloc(batch * n * sizeof(Fr), st));
    CK(cudaMemcpyAsync(raw.p, vals, batch * n * 32, cudaMemcpyHostToDevice, st));
    launch_fr_to_mont(raw.as<uint64_t>(), buf.as<Fr>(), batch * n, st);
    launch_das_fft_extension(fs->dom, buf.as<Fr>(), log2u(n), batch, fr_inv_of_u64(n), st);
    launch_fr_from_mont(buf.as<Fr>(), raw.as<uint64_t>(), batch * n, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(vals, raw.p, batch * n * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
extern "C" int b200_das_fft_extension(b200_fs* fs, uint64_t* vals, size_t n) { return b200_das_fft_extension_batch(fs, vals, n, 1); }

// ------------------------------------------------------------------------------ zero poly / recovery
// Length bookkeeping of zero_poly.go:116-217 replayed on the host: the device computes the same
// polynomial by another exact route, so the cases in which the reference *panics* (slice bounds,
// "expected larger destination length", "expected output smaller or equal to input length") have
// to be reproduced from the sizes alone.  Returns false where the reference would panic.
static bool zero_poly_sizes_ok(size_t nmiss, size_t length) {
    const size_t per_leaf = 63, per_leaf_poly = 64;
    if (nmiss <= per_leaf) return nmiss + 1 <= length;           // zero_poly.go:133 make(len > cap) panics
    size_t leaf_count = (nmiss + per_leaf - 1) / per_leaf;
    size_t n = next_pow2(leaf_count * per_leaf_poly);
    std::vector<size_t> len(leaf_count, per_leaf_poly);
    size_t nleaves = leaf_count;
    while (nleaves > 1) {
        size_t reduced_count = (nleaves + 3) / 4, leaf_size = next_pow2(len[0]);
        for (size_t i = 0; i < reduced_count; i++) {
            size_t start = i * 4, end = start + 4, out_end = end * leaf_size;
            if (out_end > n) out_end = n;
            if (start * leaf_size > out_end) return false;         // zero_poly.go:190 slice bounds
            size_t rlen = out_end - start * leaf_size;
            if (end > nleaves) end = nleaves;
            if (end > start + 1) {
                size_t deg = 0;
                for (size_t q = start; q < end; q++) deg += len[q] - 1;
                if (!is_pow2(rlen) || deg + 1 > rlen) return false;   // zero_poly.go:60-76
                rlen = deg + 1;
            }
            len[i] = rlen;
        }
        nleaves = reduced_count;
    }
    return len[0] <= length;                                       // zero_poly.go:207-209
}

// device side: zero_eval[b] and zero_poly[b] (Montgomery) for `batch` missing lists
// zero polynomial of `batch` index sets that already sit on the device (d_missing: pitch entries per set, d_nmiss: their
// sizes; max_missing: the largest size, known to the host)
static int dev_zero_poly_lists(b200_fs* fs, const uint32_t* d_missing, const uint32_t* d_nmiss, size_t max_missing, size_t pitch,
                               size_t n, size_t batch, Fr* d_zero_eval, Fr* d_zero_poly, cudaStream_t st) {
    const size_t mp = zero_poly_tree_size(max_missing);
    if (max_missing >= 256 && mp <= n && max_missing < n) {
        // large sets: coefficients through the product tree, evaluations with one forward NTT
        DevBuf ca, cb, padded, tmp;
        CKS(ca.alloc(batch * mp * sizeof(Fr), st)); CKS(cb.alloc(batch * mp * sizeof(Fr), st));
        CKS(padded.alloc(batch * 2 * mp * sizeof(Fr), st)); CKS(tmp.alloc(batch * 2 * mp * sizeof(Fr), st));
        launch_zero_poly_tree(fs->dom, n, batch, d_missing, d_nmiss, pitch, mp, ca.as<Fr>(), cb.as<Fr>(),
                              padded.as<Fr>(), tmp.as<Fr>(), d_zero_poly, st);
        CKS(check_launches());
        CKS(dev_fr_fft(fs, d_zero_poly, d_zero_eval, log2u(n), batch, false, st));
    } else {
        DevBuf partial;
        CKS(partial.alloc(batch * zero_eval_segments(max_missing) * n * sizeof(Fr), st));
        launch_zero_eval(fs->dom, n, batch, d_missing, d_nmiss, pitch, max_missing, partial.as<Fr>(), d_zero_eval, st);
        CKS(check_launches());
        CKS(dev_fr_fft(fs, d_zero_eval, d_zero_poly, log2u(n), batch, true, st));
    }
    return B200_OK;
}
static int dev_zero_poly(b200_fs* fs, const std::vector<uint32_t>& h_missing, const std::vector<uint32_t>& h_nmiss, size_t pitch,
                         size_t n, size_t batch, Fr* d_zero_eval, Fr* d_zero_poly, cudaStream_t st) {
    size_t max_missing = 0;
    for (size_t b = 0; b < batch; b++) if (h_nmiss[b] > max_missing) max_missing = h_nmiss[b];
    DevBuf miss, cnt;
    CKS(miss.alloc(h_missing.size() * 4, st)); CKS(cnt.alloc(batch * 4, st));
    CK(cudaMemcpyAsync(miss.p, h_missing.data(), h_missing.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cnt.p, h_nmiss.data(), batch * 4, cudaMemcpyHostToDevice, st));
    CKS(dev_zero_poly_lists(fs, miss.as<uint32_t>(), cnt.as<uint32_t>(), max_missing, pitch, n, batch, d_zero_eval, d_zero_poly, st));
    CK(cudaStreamSynchronize(st));   // h_missing / h_nmiss are the caller's stack vectors
    return B200_OK;
}

extern "C" int b200_zero_poly_via_multiplication(b200_fs* fs, const uint64_t* missing, size_t n_missing, size_t length,
                                                 uint64_t* zero_eval, uint64_t* zero_poly) {
    if (n_missing == 0) {                                          // zero_poly.go:117-119
        memset(zero_eval, 0, length * 32); memset(zero_poly, 0, length * 32);
        return B200_OK;
    }
    if (length > fs->max_width) return B200_ERR_TOO_SMALL;         // zero_poly.go:120-122 "domain too small"
    if (!is_pow2(length)) return B200_ERR_NOT_POW2;                // zero_poly.go:123-125
    const uint64_t stride = fs->max_width / length;
    std::vector<uint32_t> m(n_missing), cnt(1, (uint32_t)n_missing);
    for (size_t i = 0; i < n_missing; i++) {
        if (missing[i] * stride > fs->max_width) return B200_ERR_BAD_INPUT;   // root table index out of range
        m[i] = (uint32_t)(missing[i] % length);                    // w^(length stride) == w^0
    }
    if (!zero_poly_sizes_ok(n_missing, length)) return B200_ERR_BAD_INPUT;
    CK(cudaSetDevice(fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf ze, zp, raw;
    CKS(ze.alloc(length * sizeof(Fr), st)); CKS(zp.alloc(length * sizeof(Fr), st)); CKS(raw.alloc(2 * length * 32, st));
    CKS(dev_zero_poly(fs, m, cnt, n_missing, length, 1, ze.as<Fr>(), zp.as<Fr>(), st));
    launch_fr_from_mont(ze.as<Fr>(), raw.as<uint64_t>(), length, st);
    launch_fr_from_mont(zp.as<Fr>(), raw.as<uint64_t>() + 4 * length, length, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(zero_eval, raw.p, length * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(zero_poly, raw.as<uint64_t>() + 4 * length, length * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// powers of the coset shift factor 5 (recover_from_samples.go:9-40): [0] = 5^-i, [1] = 5^i, i < max_width
static int fs_shift_tables(b200_fs* fs, const Fr** inv_pows, const Fr** pows) {
    std::lock_guard<std::mutex> lk(fs->mu);
    if (!fs->d_shift[0]) {
        const uint64_t W = fs->max_width;
        std::vector<Fr> t(W);
        for (int which = 0; which < 2; which++) {
            Fr f = fr_from_u64(5);
            if (which == 0) f = fe_inv(f);
            Fr p = Fr::one();
            for (uint64_t i = 0; i < W; i++) { t[i] = p; p = fe_mul(p, f); }
            CK(cudaMalloc(&fs->d_shift[which], W * sizeof(Fr)));
            CK(cudaMemcpy(fs->d_shift[which], t.data(), W * sizeof(Fr), cudaMemcpyHostToDevice));
        }
        CK(cudaDeviceSynchronize());   // as in fs_programs
    }
    *inv_pows = fs->d_shift[0]; *pows = fs->d_shift[1];
    return B200_OK;
}

// number of zero bytes among n (present[i] == 0 <=> samples[i] == nil), eight at a time
static uint32_t count_zero_bytes(const uint8_t* p, size_t n) {
    uint32_t c = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t v;
        memcpy(&v, p + i, 8);
        const uint64_t k = 0x7f7f7f7f7f7f7f7full;
        const uint64_t t = ~(((v & k) + k) | v | k);                // 0x80 in every byte of v that is zero
        c += (uint32_t)__builtin_popcountll(t);
    }
    for (; i < n; i++) c += p[i] == 0;
    return c;
}

// recover_from_samples.go:42-109 for a batch.  The host only counts the missing samples (every size check of the reference
// depends on the counts alone); the index lists are compacted on the device from the presence mask -- built on the host they
// cost more than the whole device pipeline.  On an error return the contents of `out` are unspecified.
extern "C" int b200_recover_poly_from_samples_batch(b200_fs* fs, const uint64_t* samples, const uint8_t* present, size_t n,
                                                    size_t batch, uint64_t* out) {
    if (batch == 0) return B200_OK;
    if (batch > 65535) return B200_ERR_TOO_LARGE;                  // the batch rides on grid.y / grid.z
    std::vector<uint32_t> cnt(batch);
    size_t max_missing = 0;
    for (size_t b = 0; b < batch; b++) {
        cnt[b] = count_zero_bytes(present + b * n, n);
        if (cnt[b] > max_missing) max_missing = cnt[b];
    }
    for (size_t b = 0; b < batch; b++) {
        // nothing missing: zeroPolyFn returns all zeros and the sanity loop panics (recover_from_samples.go:54-58)
        if (cnt[b] == 0) return n == 0 ? B200_OK : B200_ERR_ZERO_EVAL;
    }
    if (n > fs->max_width) return B200_ERR_TOO_SMALL;              // zero_poly.go:120-122
    if (!is_pow2(n)) return B200_ERR_NOT_POW2;                     // zero_poly.go:123-125
    for (size_t b = 0; b < batch; b++) if (!zero_poly_sizes_ok(cnt[b], n)) return B200_ERR_BAD_INPUT;
    CK(cudaSetDevice(fs->device));
    const Fr *shift_inv, *shift_fwd;
    CKS(fs_shift_tables(fs, &shift_inv, &shift_fwd));
    const unsigned logn = log2u(n);
    const size_t total = batch * n;
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf raw, s, ze, zp, a, c, pres, flags, miss, dcnt;
    CKS(raw.alloc(total * 32, st)); CKS(s.alloc(total * sizeof(Fr), st)); CKS(ze.alloc(total * sizeof(Fr), st));
    CKS(zp.alloc(total * sizeof(Fr), st)); CKS(a.alloc(total * sizeof(Fr), st)); CKS(c.alloc(total * sizeof(Fr), st));
    CKS(pres.alloc(total, st)); CKS(flags.alloc(batch * 4, st)); CKS(miss.alloc(total * 4, st)); CKS(dcnt.alloc(batch * 4, st));
    CK(cudaMemcpyAsync(pres.p, present, total, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(raw.p, samples, total * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(flags.p, 0, batch * 4, st));
    launch_missing_lists(pres.as<uint8_t>(), n, batch, miss.as<uint32_t>(), n, dcnt.as<uint32_t>(), st);   // :45-50
    launch_fr_to_mont(raw.as<uint64_t>(), s.as<Fr>(), total, st);
    CKS(dev_zero_poly_lists(fs, miss.as<uint32_t>(), dcnt.as<uint32_t>(), max_missing, n, n, batch, ze.as<Fr>(), zp.as<Fr>(), st));
    launch_fr_mul_masked(a.as<Fr>(), s.as<Fr>(), ze.as<Fr>(), pres.as<uint8_t>(), total, st);   // E = samples (.) zeroEval
    CKS(dev_fr_fft(fs, a.as<Fr>(), a.as<Fr>(), logn, batch, true, st));                            // polyWithZero
    launch_fr_mul_table(a.as<Fr>(), shift_inv, n, batch, st);                                      // ShiftPoly
    launch_fr_mul_table(zp.as<Fr>(), shift_inv, n, batch, st);
    CKS(dev_fr_fft(fs, a.as<Fr>(), a.as<Fr>(), logn, batch, false, st));                           // evalShiftedPolyWithZero
    CKS(dev_fr_fft(fs, zp.as<Fr>(), c.as<Fr>(), logn, batch, false, st));                          // evalShiftedZeroPoly
    launch_fr_div(a.as<Fr>(), c.as<Fr>(), total, st);                                              // :89-91
    CKS(dev_fr_fft(fs, a.as<Fr>(), a.as<Fr>(), logn, batch, true, st));
    launch_fr_mul_table(a.as<Fr>(), shift_fwd, n, batch, st);                                      // UnshiftPoly
    CKS(dev_fr_fft(fs, a.as<Fr>(), a.as<Fr>(), logn, batch, false, st));                           // reconstructedData
    launch_recover_check(a.as<Fr>(), s.as<Fr>(), ze.as<Fr>(), pres.as<uint8_t>(), n, batch, flags.as<uint32_t>(), st);
    launch_fr_from_mont(a.as<Fr>(), raw.as<uint64_t>(), total, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out, raw.p, total * 32, cudaMemcpyDeviceToHost, st));
    std::vector<uint32_t> h_flags(batch);
    CK(cudaMemcpyAsync(h_flags.data(), flags.p, batch * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (size_t b = 0; b < batch; b++) {
        if (h_flags[b] & 2) return B200_ERR_ZERO_EVAL;             // recover_from_samples.go:54-58 (panic)
        if (h_flags[b] & 1) return B200_ERR_RECOVERY;              // recover_from_samples.go:103-107 (error)
    }
    return B200_OK;
}
extern "C" int b200_recover_poly_from_samples(b200_fs* fs, const uint64_t* s, const uint8_t* p, size_t n, uint64_t* out) {
    return b200_recover_poly_from_samples_batch(fs, s, p, n, 1, out);
}

// ------------------------------------------------------------------------------ G1 FFT
// In-place transform of `batch` vectors of n = 2^logn points (element i of blob b at
// data[b * bstride + i * estride]).  dif: natural order in, bit-reversed out; otherwise (DIT)
// bit-reversed in, natural out.  No 1/n scaling here.
// skip_dif: leading decimation-in-frequency stages already done elsewhere (dev_fk20 folds two into ToeplitzPart2)
// One stage over `batch` transforms: whole-warp batches share the twiddle across the lanes of a warp (sparse programs);
// for fewer transforms the lanes run across the blocks of the stage while it has at least 32 of them (same sparse
// programs), and per-lane fixed-window programs are left for the few last (DIT) / first (DIF) stages.
struct StagePrograms { const ScalarProgram* per_lane; const ScalarProgram* shared; };
static int fs_stage_programs(b200_fs* fs, int inverse, size_t batch, StagePrograms* sp) {
    sp->per_lane = sp->shared = nullptr;
    CKS(fs_programs(fs, inverse, 1, &sp->shared));
    if (program_mode_for_batch(batch) == 0) CKS(fs_programs(fs, inverse, 0, &sp->per_lane));
    return B200_OK;
}
static void launch_stage_auto(const StagePrograms& sp, G1J* data, size_t n_half, size_t batch, size_t m, size_t estride, size_t bstride,
                              bool dif, size_t prog_stride, cudaStream_t st) {
    const size_t share = g1_stage_uses_quads(n_half, batch) ? 8 : 32;      // units of a warp that must hold the same twiddle
    if (program_mode_for_batch(batch) == 1) launch_g1_fft_stage(data, n_half, batch, m, estride, bstride, dif, sp.shared, prog_stride, st, 0);
    else if ((n_half / m) % share == 0) launch_g1_fft_stage(data, n_half, batch, m, estride, bstride, dif, sp.shared, prog_stride, st, 1);
    else launch_g1_fft_stage(data, n_half, batch, m, estride, bstride, dif, sp.per_lane, prog_stride, st, 0);
}
static int dev_g1_fft_stages(b200_fs* fs, G1J* data, unsigned logn, size_t batch, size_t estride, size_t bstride,
                             bool inverse, bool dif, cudaStream_t st, unsigned skip_dif = 0) {
    if (logn == 0) return B200_OK;
    StagePrograms sp;
    CKS(fs_stage_programs(fs, inverse ? 1 : 0, batch, &sp));
    const size_t n = (size_t)1 << logn, halfw = fs->max_width / 2;
    if (dif) {
        for (size_t m = (n / 2) >> skip_dif; m >= 1; m >>= 1)
            launch_stage_auto(sp, data, n / 2, batch, m, estride, bstride, true, halfw / m, st);
    } else {
        for (size_t m = 1; m <= n / 2; m <<= 1)
            launch_stage_auto(sp, data, n / 2, batch, m, estride, bstride, false, halfw / m, st);
    }
    return check_launches();
}

// FFTG1 of `batch` host vectors; only the first out_count <= n outputs of every transform are returned
static int host_fft_g1(b200_fs* fs, const uint64_t* vals, size_t n, size_t batch, int inverse, uint64_t* out, size_t out_count) {
    if (n > fs->max_width) return B200_ERR_TOO_LARGE;     // fft_g1.go:60-62
    if (n == 0) return B200_ERR_BAD_INPUT;                // bls.IsPowerOfTwo(0) holds, then fft_g1.go:78,88 divides by n: panic
    if (!is_pow2(n)) return B200_ERR_NOT_POW2;            // fft_g1.go:63-65
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fs->device));
    const unsigned logn = log2u(n);
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf raw, buf;
    CKS(raw.alloc(batch * n * 144, st));
    CKS(buf.alloc(batch * n * sizeof(G1J), st));
    CK(cudaMemcpyAsync(raw.p, vals, batch * n * 144, cudaMemcpyHostToDevice, st));
    launch_g1_from_abi(raw.as<uint64_t>(), buf.as<G1J>(), batch * n, st);
    CKS(dev_g1_fft_stages(fs, buf.as<G1J>(), logn, batch, 1, n, inverse != 0, true, st));
    DevBuf prog;
    if (inverse) {   // fft_g1.go:81-84: every output times n^-1
        ScalarProgram sp;
        make_scalar_program(&sp, fe_from_mont(fr_inv_of_u64(n)), 1);
        CKS(prog.alloc(sizeof(ScalarProgram), st));
        CK(cudaMemcpyAsync(prog.p, &sp, sizeof sp, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));   // sp lives on this stack frame
        launch_g1_mul_programs(buf.as<G1J>(), n, batch, 1, n, prog.as<ScalarProgram>(), 0, 0, logn, st);
    }
    // natural-order output i sits at bit-reversed position rev(i) after the decimation-in-frequency stages
    launch_g1_to_abi(buf.as<G1J>(), raw.as<uint64_t>(), n, batch, 1, n, 1, logn, st);
    CKS(check_launches());
    CK(cudaMemcpy2DAsync(out, out_count * 144, raw.p, n * 144, out_count * 144, batch, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
extern "C" int b200_fft_g1_batch(b200_fs* fs, const uint64_t* vals, size_t n, size_t batch, int inverse, uint64_t* out) {
    return host_fft_g1(fs, vals, n, batch, inverse, out, n);
}
extern "C" int b200_fft_g1(b200_fs* fs, const uint64_t* vals, size_t n, int inverse, uint64_t* out) {
    return host_fft_g1(fs, vals, n, 1, inverse, out, n);
}
// das_extension.go:7-84 over G1 -- the "G1 version of the DAS extension FFT" of the TODO at fk20_multi.go:96: given the
// even-index evaluations of a polynomial of degree < n over G1, the odd-index ones.  Same butterfly network as the Fr form
// (kernels_fr.cu: descent = inverse DIF stages down to pairs, a middle butterfly with the root Exp[n / 2], ascent = DIT
// stages with the odd roots Exp[(1 + 2 i) n / L], finally 1 / n), run with the G1 stage kernel; roots are indexed on
// the full domain with stride 1 whatever n is, like the reference.  In place, n a power of two >= 2, 2 n <= MaxWidth.
// (The FK20 pipelines do not use it: they hold the COEFFICIENTS h, for which the zero-padded transform costs one twist
// plus two half-size transforms -- fewer multiplications than a transform plus this extension.  DESIGN.md section 3.)
extern "C" int b200_das_fft_extension_g1(b200_fs* fs, uint64_t* vals, size_t n) {
    if (n * 2 > fs->max_width) return B200_ERR_TOO_SMALL;   // das_extension.go:72-74
    if (n < 2 || !is_pow2(n)) return B200_ERR_BAD_INPUT;    // das_extension.go:22-24 "bad usage"
    CK(cudaSetDevice(fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    const unsigned logn = log2u(n);
    DevBuf raw, buf, prog;
    CKS(raw.alloc(n * 144, st)); CKS(buf.alloc(n * sizeof(G1J), st));
    CK(cudaMemcpyAsync(raw.p, vals, n * 144, cudaMemcpyHostToDevice, st));
    launch_g1_from_abi(raw.as<uint64_t>(), buf.as<G1J>(), n, st);
    StagePrograms spi, spf;
    CKS(fs_stage_programs(fs, 1, 1, &spi));
    CKS(fs_stage_programs(fs, 0, 1, &spf));
    G1J* d = buf.as<G1J>();
    for (size_t m = n / 2; m >= 2; m >>= 1)                  // descent: (a0 + a1, (a0 - a1) Rev[i n / m]), i < m
        launch_stage_auto(spi, d, n / 2, 1, m, 1, n, true, n / m, st);
    {                                                        // pairs: x = a0 + a1, t = (a0 - a1) Exp[n / 2]; (x + t, x - t)
        StagePrograms mid = {spf.per_lane + n / 2, spf.shared + n / 2};
        launch_stage_auto(mid, d, n / 2, 1, 1, 1, n, true, 0, st);
        launch_stage_auto(spf, d, n / 2, 1, 1, 1, n, false, 0, st);
    }
    for (size_t m = 2; m <= n / 2; m <<= 1) {                // ascent: (a0 + a1 Exp[(1 + 2 i) n / (2 m)], a0 - ..), i < m
        StagePrograms odd = {spf.per_lane + n / (2 * m), spf.shared + n / (2 * m)};
        launch_stage_auto(odd, d, n / 2, 1, m, 1, n, false, n / m, st);
    }
    ScalarProgram sp;
    make_scalar_program(&sp, fe_from_mont(fr_inv_of_u64(n)), 1);
    CKS(prog.alloc(sizeof(ScalarProgram), st));
    CK(cudaMemcpyAsync(prog.p, &sp, sizeof sp, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                           // sp lives on this stack frame
    launch_g1_mul_programs(d, n, 1, 1, n, prog.as<ScalarProgram>(), 0, 0, logn, st);
    launch_g1_to_abi(d, raw.as<uint64_t>(), n, 1, 1, n, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(vals, raw.p, n * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// fk20_single.go:80-87 ToeplitzPart3: inverse FFTG1, first half of the result (n2 / 2 points)
extern "C" int b200_toeplitz_part3(b200_fs* fs, const uint64_t* h_ext_fft, size_t n2, uint64_t* out) {
    return host_fft_g1(fs, h_ext_fft, n2, 1, 1, out, n2 / 2);
}

// ------------------------------------------------------------------------------ LinCombG1
// sum_i k[b][i] * pts[i] for each blob b; result of blob b is left at work[b * n].  d_k canonical or Montgomery
// ([batch][n]).  Three routes: fixed-base window table (settings-owned bases), Pippenger bucket MSM (variable bases,
// kernels_msm.cu), and for a handful of terms per-term windowed multiplication + fold tree.
static const size_t kMsmMinTerms = 32;
static int dev_lincomb(const G1J* d_pts, size_t pts_bstride, const Fr* d_k, int k_is_mont, G1J* work, size_t n,
                       size_t batch, cudaStream_t st, const G1A* fb_table = nullptr, int fb_w = 8) {
    // The bucket method pays per MSM a latency-bound tail of ~1.3 ms (bucket reduction + 120 doublings), so several small MSMs
    // in one call are cheaper as one launch of per-term products (72 ns per term at full occupancy) + fold tree.
    if (!fb_table && n >= kMsmMinTerms && (batch == 1 || n >= 16384)) {
        DevBuf ws;
        CKS(ws.alloc(msm_workspace_bytes(n), st));
        for (size_t b = 0; b < batch; b++)
            launch_g1_msm(d_pts + b * pts_bstride, d_k + b * n, k_is_mont, n, ws.p, work + b * n, st);
        return check_launches();
    }
    if (fb_table) launch_g1_mul_fixed_base(fb_table, fb_w, d_k, k_is_mont, work, n, n, batch, st);
    else launch_g1_mul_var(d_pts, pts_bstride, d_k, k_is_mont, work, n, n, batch, st);
    for (size_t cnt = n; cnt > 1;) {
        size_t half = (cnt + 1) / 2;
        launch_g1_fold(work, n, half, cnt, batch, st);
        cnt = half;
    }
    return check_launches();
}

extern "C" int b200_g1_lincomb(const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t* out) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    if (n == 0) { memset(out, 0, 144); return B200_OK; }     // bls/bls_test.go:69-77: empty sum is infinity
    CK(cudaSetDevice(g_device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf raw, pts, k, work;
    CKS(raw.alloc(n * 144, st)); CKS(pts.alloc(n * sizeof(G1J), st)); CKS(k.alloc(n * 32, st)); CKS(work.alloc(n * sizeof(G1J), st));
    CK(cudaMemcpyAsync(raw.p, points, n * 144, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(k.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
    launch_g1_from_abi(raw.as<uint64_t>(), pts.as<G1J>(), n, st);
    CKS(dev_lincomb(pts.as<G1J>(), 0, k.as<Fr>(), 0, work.as<G1J>(), n, 1, st));
    launch_g1_to_abi(work.as<G1J>(), raw.as<uint64_t>(), 1, 1, 1, n, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out, raw.p, 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

extern "C" int b200_g1_mul_many(const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t* out) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    if (n == 0) return B200_OK;
    CK(cudaSetDevice(g_device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf raw, pts, k;
    CKS(raw.alloc(n * 144, st)); CKS(pts.alloc(n * sizeof(G1J), st)); CKS(k.alloc(n * 32, st));
    CK(cudaMemcpyAsync(raw.p, points, n * 144, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(k.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
    launch_g1_from_abi(raw.as<uint64_t>(), pts.as<G1J>(), n, st);
    // batch = n, one element each: per-lane scalars
    launch_g1_mul_var(pts.as<G1J>(), 1, k.as<Fr>(), 0, pts.as<G1J>(), 1, 1, n, st);
    launch_g1_to_abi(pts.as<G1J>(), raw.as<uint64_t>(), n, 1, 1, n, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out, raw.p, n * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// fk20_single.go:59-77 ToeplitzPart2 with caller-held points: hExtFFT[i] = FFT(toeplitzCoeffs)[i] * xExtFFT[i].
// (The FK20 entry points run a fused form over the settings' resident xExtFFT; this is the stand-alone method.)
extern "C" int b200_toeplitz_part2(b200_fs* fs, const uint64_t* toeplitz_coeffs, size_t n_coeffs, const uint64_t* x_ext_fft, size_t n_points,
                                   uint64_t* h_ext_fft) {
    if (n_coeffs != n_points) return B200_ERR_LEN_MISMATCH;          // fk20_single.go:60-62 (panic)
    if (n_coeffs > fs->max_width) return B200_ERR_TOO_LARGE;         // FFT error -> panic (:63-66)
    const size_t np = next_pow2(n_coeffs);
    if (np != n_points) return B200_ERR_LEN_MISMATCH;                // FFT pads (fft_fr.go:60); xExtFFT[i] would run out of range
    CK(cudaSetDevice(fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    const unsigned logn = log2u(np);
    DevBuf raw, c, pts;
    CKS(raw.alloc(np * 144, st)); CKS(c.alloc(np * sizeof(Fr), st)); CKS(pts.alloc(np * sizeof(G1J), st));
    CK(cudaMemcpyAsync(raw.p, toeplitz_coeffs, np * 32, cudaMemcpyHostToDevice, st));
    launch_fr_to_mont(raw.as<uint64_t>(), c.as<Fr>(), np, st);
    CKS(dev_fr_fft(fs, c.as<Fr>(), c.as<Fr>(), logn, 1, false, st));
    CK(cudaMemcpyAsync(raw.p, x_ext_fft, np * 144, cudaMemcpyHostToDevice, st));
    launch_g1_from_abi(raw.as<uint64_t>(), pts.as<G1J>(), np, st);
    launch_g1_mul_var(pts.as<G1J>(), 1, c.as<Fr>(), 1, pts.as<G1J>(), 1, 1, np, st);   // per-lane scalars
    launch_g1_to_abi(pts.as<G1J>(), raw.as<uint64_t>(), np, 1, 1, np, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(h_ext_fft, raw.p, np * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// bls/bls_kilic.go:118-121 FromCompressedG1 over an array, on the device (square root + subgroup check per point).
// ok (may be NULL) receives 1 per accepted encoding; any rejected encoding makes the call return B200_ERR_BAD_INPUT
// (the reference returns an error for that element), with all-zero output for the rejected points.
extern "C" int b200_g1_from_compressed_batch(const uint8_t* in48, size_t n, uint64_t* out, uint8_t* ok) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    if (n == 0) return B200_OK;
    CK(cudaSetDevice(g_device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf in, res, stt;
    CKS(in.alloc(n * 48, st)); CKS(res.alloc(n * 144, st)); CKS(stt.alloc(n * 4, st));
    CK(cudaMemcpyAsync(in.p, in48, n * 48, cudaMemcpyHostToDevice, st));
    launch_g1_decompress(in.as<uint8_t>(), res.as<uint64_t>(), stt.as<uint32_t>(), n, st);
    CKS(check_launches());
    std::vector<uint32_t> h(n);
    CK(cudaMemcpyAsync(out, res.p, n * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h.data(), stt.p, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int rc = B200_OK;
    for (size_t i = 0; i < n; i++) {
        if (ok) ok[i] = h[i] == 0;
        if (h[i]) rc = B200_ERR_BAD_INPUT;
    }
    return rc;
}

// bls/globals.go:106-153 EvaluatePolyInEvaluationForm for a batch: polys[b] = n evaluations on the settings' domain of
// order n (natural order, or reverse bit order like eth's DomainFr), xs[b] the evaluation point; y[b] canonical.
// A point inside the domain gives 0, as in the reference (BatchInvModFr leaves the zero denominator, x^n - 1 = 0).
extern "C" int b200_evaluate_poly_in_evaluation_form_batch(b200_fs* fs, const uint64_t* polys, const uint64_t* xs, size_t n, size_t batch,
                                                           int reverse_bit_order, uint64_t* y) {
    if (n > fs->max_width || n == 0 || !is_pow2(n)) return B200_ERR_LEN_MISMATCH;   // bls/globals.go:107-109 (panic): no such roots
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    const unsigned logn = log2u(n);
    DevBuf f, dx, inv, part, ym, yc;
    CKS(f.alloc(batch * n * 32, st)); CKS(dx.alloc(batch * 32, st)); CKS(inv.alloc(batch * n * 32, st));
    CKS(part.alloc(batch * (n / 16 + 1) * 32, st)); CKS(ym.alloc(batch * 32, st)); CKS(yc.alloc(batch * 32, st));
    CK(cudaMemcpyAsync(f.p, polys, batch * n * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dx.p, xs, batch * 32, cudaMemcpyHostToDevice, st));
    launch_eval_form_quotient(fs->dom, f.as<uint64_t>(), dx.as<uint64_t>(), logn, batch, reverse_bit_order ? 1 : 0, fr_inv_of_u64(n), inv.as<Fr>(),
                              part.as<Fr>(), ym.as<Fr>(), yc.as<uint64_t>(), nullptr, nullptr, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(y, yc.p, batch * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// kzg_single_proofs.go:57-75 CheckProofSingle, G1 side for a batch: out[i] = commitment[i] - y[i] G (:63-67).
// The caller finishes with its pairing backend: e(out[i], [1]_2) == e(proof[i], [s - x[i]]_2).
extern "C" int b200_check_proof_single_g1_batch(const uint64_t* commitments, const uint64_t* ys, size_t batch, uint64_t* out) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(g_device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    G1J gen = g1_generator();
    DevBuf raw, c, k, dg, yg;
    CKS(raw.alloc(batch * 144, st)); CKS(c.alloc(batch * sizeof(G1J), st)); CKS(k.alloc(batch * 32, st));
    CKS(dg.alloc(sizeof(G1J), st)); CKS(yg.alloc(batch * sizeof(G1J), st));
    CK(cudaMemcpyAsync(raw.p, commitments, batch * 144, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(k.p, ys, batch * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dg.p, &gen, sizeof gen, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                                   // gen lives on this stack frame
    launch_g1_from_abi(raw.as<uint64_t>(), c.as<G1J>(), batch, st);
    launch_g1_mul_var(dg.as<G1J>(), 0, k.as<Fr>(), 0, yg.as<G1J>(), 1, 1, batch, st);   // y[i] G: one shared base
    launch_g1_sub_arrays(c.as<G1J>(), yg.as<G1J>(), 1, batch, st);
    launch_g1_to_abi(c.as<G1J>(), raw.as<uint64_t>(), batch, 1, 1, batch, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out, raw.p, batch * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// ------------------------------------------------------------------------------ KZGSettings
struct b200_ks {
    int device = 0;
    b200_fs* fs = nullptr;
    size_t n_g1 = 0;
    G1J* d_secret_g1 = nullptr;   // Montgomery Jacobian
    std::mutex mu;
    G1A* d_fb_table = nullptr;    // fixed-base window table of SecretG1[:fb_n] (built on first commit)
    int fb_w = 8;                 // its window bits
    size_t fb_n = 0;
    std::vector<G1A*> retired;    // smaller tables superseded by d_fb_table (freed with the settings)
    std::vector<uint64_t> h_secret_g2;   // SecretG2 (kzg.go:16) as handed in (36 x u64 per point): only the verification entry points read it
    size_t n_secret_g2() const { return h_secret_g2.size() / 36; }
};

// Fixed-base window tables.  Window bits: 8 by default (384 KiB per base: 1.5 GiB for the 4096 commitment bases,
// 3 GiB for the 8192 bases of the n = 4096 FK20 settings); wider windows trade HBM for additions and are opt-in
// (b200_set_fixed_base_window / B200_FB_WINDOW = 10 or 12: 1.3 / 4.1 MiB per base); 4 bits (48 KiB per base) when
// the preferred table does not fit (config 5: 2.1 M bases).
static std::atomic<int> g_fb_window{0};   // 0: not set -> environment -> 8
static int preferred_fb_window() {
    int w = g_fb_window.load();
    if (w == 0) {
        const char* e = getenv("B200_FB_WINDOW");
        w = e ? atoi(e) : 8;
        if (w != 4 && w != 8 && w != 10 && w != 12) w = 8;
        g_fb_window = w;
    }
    return w;
}
extern "C" int b200_set_fixed_base_window(int bits) {
    if (bits != 4 && bits != 8 && bits != 10 && bits != 12) return B200_ERR_BAD_INPUT;
    g_fb_window = bits;
    return B200_OK;
}
static const size_t kFixedBaseBudgetWide = (size_t)40 << 30;  // 10 / 12-bit windows up to here and 40 % of the free HBM
static const size_t kFixedBaseBudget8 = (size_t)8 << 30;      // 8-bit windows (384 KiB per base) up to here
static const size_t kFixedBaseBudget4 = (size_t)128 << 30;    // else 4-bit windows (48 KiB per base) up to here / 70 % of the free HBM

// Window table over d_pts[0..n): *out_w = window bits, *out_table = nullptr when nothing fits (the callers
// then use the bucket MSM / generic windowed multiplication).  Running out of memory is not an error either.
static int build_fixed_base(const G1J* d_pts, size_t n, G1A** out_table, int* out_w, cudaStream_t st) {
    *out_table = nullptr; *out_w = 8;
    if (n == 0) return B200_OK;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return B200_OK; }
    // the preferred width first, then the narrower ones; an allocation failure moves on to the next choice
    G1A* table = nullptr;
    G1J* tmp = nullptr;
    int W = 0;
    const int pref = preferred_fb_window();
    const int choices[4] = {12, 10, 8, 4};
    for (int c = 0; c < 4 && !table; c++) {
        const int w = choices[c];
        if (w > pref) continue;
        const size_t tb = fixed_base_table_bytes(n, w), need = tb + fixed_base_tmp_bytes(n, w);
        bool fits = w >= 10 ? (tb <= kFixedBaseBudgetWide && need <= free_b / 10 * 4)
                  : w == 8  ? (tb <= kFixedBaseBudget8 && need <= free_b / 10 * 7)
                            : (tb <= kFixedBaseBudget4 && need <= free_b / 10 * 7);
        if (!fits) continue;
        if (cudaMalloc(&table, tb) != cudaSuccess) { cudaGetLastError(); table = nullptr; continue; }
        if (cudaMalloc(&tmp, fixed_base_tmp_bytes(n, w)) != cudaSuccess) { cudaGetLastError(); cudaFree(table); table = nullptr; tmp = nullptr; continue; }
        W = w;
    }
    if (!table) return B200_OK;
    launch_fixed_base_table(d_pts, n, tmp, table, W, st);
    int rc = check_launches();
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { g_cuda_err = "fixed-base table build"; rc = B200_ERR_CUDA; }
    cudaFree(tmp);
    if (rc) { cudaFree(table); return rc; }
    *out_table = table; *out_w = W;
    return B200_OK;
}
// Table covering SecretG1[:n] (null when over budget) and its window width, handed out together under the lock.
// Tables only ever grow: a table superseded by a larger one stays allocated until the settings are freed, because
// a concurrent call on another thread may still have kernels in flight that read it.
static int ks_fixed_base(b200_ks* ks, size_t n, const G1A** table, int* w, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(ks->mu);
    if (ks->fb_n < n) {
        G1A* t = nullptr;
        int tw = 8;
        CKS(build_fixed_base(ks->d_secret_g1, n, &t, &tw, st));
        if (!t) { *table = nullptr; *w = 8; return B200_OK; }
        if (ks->d_fb_table) ks->retired.push_back(ks->d_fb_table);
        ks->fb_w = tw;
        ks->d_fb_table = t;
        ks->fb_n = n;
    }
    *table = ks->d_fb_table; *w = ks->fb_w;
    return B200_OK;
}

extern "C" int b200_kzg_settings_new(b200_fs* fs, const uint64_t* secret_g1, size_t n_g1, size_t n_g2, b200_ks** out) {
    *out = nullptr;
    if (n_g1 != n_g2) return B200_ERR_LEN_MISMATCH;      // kzg.go:22-24
    if (n_g1 < fs->max_width) return B200_ERR_TOO_SMALL;  // kzg.go:25-27
    CK(cudaSetDevice(fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    b200_ks* ks = new (std::nothrow) b200_ks();
    if (!ks) return B200_ERR_CUDA;
    ks->fs = fs; ks->n_g1 = n_g1; ks->device = fs->device;
    DevBuf raw;
    int rc = raw.alloc(n_g1 * 144, st);
    if (!rc && cudaMalloc(&ks->d_secret_g1, (n_g1 ? n_g1 : 1) * sizeof(G1J)) != cudaSuccess) rc = B200_ERR_CUDA;
    if (rc) { delete ks; return rc; }
    if (n_g1) {
        cudaMemcpyAsync(raw.p, secret_g1, n_g1 * 144, cudaMemcpyHostToDevice, st);
        launch_g1_from_abi(raw.as<uint64_t>(), ks->d_secret_g1, n_g1, st);
    }
    rc = check_launches();
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = B200_ERR_CUDA;
    if (rc) { b200_kzg_settings_free(ks); return rc; }
    *out = ks;
    return B200_OK;
}
extern "C" void b200_kzg_settings_free(b200_ks* ks) {
    if (!ks) return;
    cudaSetDevice(ks->device);
    cudaFree(ks->d_secret_g1);
    cudaFree(ks->d_fb_table);
    for (G1A* t : ks->retired) cudaFree(t);
    delete ks;
}

extern "C" int b200_commit_to_poly_batch(b200_ks* ks, const uint64_t* coeffs, size_t n, size_t batch, uint64_t* out) {
    if (n > ks->n_g1) return B200_ERR_LEN_MISMATCH;   // SecretG1[:n] would panic (kzg_single_proofs.go:18)
    if (batch == 0) return B200_OK;
    if (n == 0) { memset(out, 0, batch * 144); return B200_OK; }
    CK(cudaSetDevice(ks->fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf k, work, res;
    CKS(k.alloc(batch * n * 32, st)); CKS(work.alloc(batch * n * sizeof(G1J), st)); CKS(res.alloc(batch * 144, st));
    CK(cudaMemcpyAsync(k.p, coeffs, batch * n * 32, cudaMemcpyHostToDevice, st));
    const G1A* fb = nullptr;
    int fbw = 8;
    if (batch * n >= 4096) CKS(ks_fixed_base(ks, n, &fb, &fbw, st));     // worth a table only for real workloads
    CKS(dev_lincomb(ks->d_secret_g1, 0, k.as<Fr>(), 0, work.as<G1J>(), n, batch, st, fb, fbw));
    launch_g1_to_abi(work.as<G1J>(), res.as<uint64_t>(), 1, batch, 1, n, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out, res.p, batch * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
extern "C" int b200_commit_to_poly(b200_ks* ks, const uint64_t* coeffs, size_t n, uint64_t* out) {
    return b200_commit_to_poly_batch(ks, coeffs, n, 1, out);
}

// kzg_multi_proofs.go:47-88 CheckProofMulti for a batch of samples, everything but the G2 arithmetic and the pairing:
// interpolation of the n = len(ys) values on the coset x <w_n> (inverse FFT, coefficient i divided by x^i, :50-70),
// [interpolation_polynomial(s)]_1 as an MSM over SecretG1[:n] (:78) and out_g1[b] = commitment[b] - that (:80-81);
// x_pow_n[b] = x[b]^n for the caller's [s^n - x^n]_2 (:71-76).  The caller finishes with
// e(out_g1[b], [1]_2) == e(proof[b], [s^n - x^n]_2).
extern "C" int b200_check_proof_multi_g1_batch(b200_ks* ks, const uint64_t* commitments, const uint64_t* xs, const uint64_t* ys, size_t n,
                                               size_t batch, uint64_t* out_g1, uint64_t* x_pow_n) {
    b200_fs* fs = ks->fs;
    if (n > fs->max_width) return B200_ERR_TOO_LARGE;        // FFT error -> panic("ys is bad") :52-54
    if (n == 0 || !is_pow2(n)) return B200_ERR_NOT_POW2;     // "The ys must have a power of 2 length" (:46)
    if (n >= ks->n_g1) return B200_ERR_LEN_MISMATCH;         // SecretG2[len(ys)] / SecretG1[:n] out of range (:76,78)
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    const unsigned logn = log2u(n);
    DevBuf raw, v, dx, xn, c, work;
    CKS(raw.alloc(batch * (n * 32 > 144 ? n * 32 : 144), st)); CKS(v.alloc(batch * n * sizeof(Fr), st)); CKS(dx.alloc(batch * 32, st));
    CKS(xn.alloc(batch * 32, st)); CKS(c.alloc(batch * sizeof(G1J), st)); CKS(work.alloc(batch * n * sizeof(G1J), st));
    CK(cudaMemcpyAsync(raw.p, ys, batch * n * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dx.p, xs, batch * 32, cudaMemcpyHostToDevice, st));
    launch_fr_to_mont(raw.as<uint64_t>(), v.as<Fr>(), batch * n, st);
    CKS(dev_fr_fft(fs, v.as<Fr>(), v.as<Fr>(), logn, batch, true, st));
    launch_unscale_coset(v.as<Fr>(), dx.as<uint64_t>(), logn, batch, xn.as<uint64_t>(), st);      // canonical coefficients now
    const G1A* fb = nullptr;
    int fbw = 8;
    if (batch * n >= 4096) CKS(ks_fixed_base(ks, n, &fb, &fbw, st));
    CKS(dev_lincomb(ks->d_secret_g1, 0, v.as<Fr>(), 0, work.as<G1J>(), n, batch, st, fb, fbw));
    CK(cudaMemcpyAsync(raw.p, commitments, batch * 144, cudaMemcpyHostToDevice, st));
    launch_g1_from_abi(raw.as<uint64_t>(), c.as<G1J>(), batch, st);
    launch_g1_sub_arrays(c.as<G1J>(), work.as<G1J>(), n, batch, st);                                  // blob b's sum sits at work[b n]
    launch_g1_to_abi(c.as<G1J>(), raw.as<uint64_t>(), batch, 1, 1, batch, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out_g1, raw.p, batch * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(x_pow_n, xn.p, batch * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// KZGSettings.SecretG2 (kzg.go:14-16).  Only CheckProofSingle / CheckProofMulti read it (SecretG2[1], SecretG2[len(ys)]), on the
// host; call once, before the handle is shared between threads.  n may be smaller than the G1 half (e.g. only the entries the
// verifier needs); every point must be on the twist.
extern "C" int b200_kzg_settings_set_secret_g2(b200_ks* ks, const uint64_t* secret_g2, size_t n) {
    std::vector<uint64_t> pts(secret_g2, secret_g2 + 36 * n);
    std::lock_guard<std::mutex> lk(ks->mu);
    ks->h_secret_g2.swap(pts);
    return B200_OK;
}
// SecretG2[i], validated when it is read (a setup can hold millions of points of which a verifier reads one or two)
static int ks_secret_g2(const b200_ks* ks, size_t i, G2J* out) {
    if (i >= ks->n_secret_g2()) return B200_ERR_TOO_SMALL;
    const uint64_t* p = ks->h_secret_g2.data() + 36 * i;
    if (!abi_coords_canonical(p, 6)) return B200_ERR_BAD_INPUT;
    *out = g2_from_abi(p);
    return g2_on_curve(*out) ? B200_OK : B200_ERR_BAD_INPUT;
}

// ok[i] = e(lhs[i], [1]_2) == e(proofs[i], t2 - cs[i] [1]_2) on the host cores: one G2 scalar multiplication and one pairing
// check per item, both inside the worker threads
static int pairing_checks(const uint64_t* lhs, const uint64_t* proofs, const G2J& t2, const uint64_t* cs, size_t batch, uint8_t* ok) {
    for (size_t i = 0; i < batch; i++) {
        if (!fr_canon_valid(cs + 4 * i)) return B200_ERR_BAD_INPUT;
        if (!abi_coords_canonical(proofs + 18 * i, 3) || !g1_on_curve(g1_from_abi_h(proofs + 18 * i))) return B200_ERR_BAD_INPUT;
    }
    const G2J gen = g2_generator();
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > batch) nt = (unsigned)batch;
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < batch;) {
            const G2J rhs = g2_sub(t2, g2_mul(gen, fr_load_canon(cs + 4 * i)));
            ok[i] = pairings_verify(g1_from_abi_h(lhs + 18 * i), gen, g1_from_abi_h(proofs + 18 * i), rhs) ? 1 : 0;
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return B200_OK;
}

// kzg_single_proofs.go:57-75 CheckProofSingle for `batch` (commitment, proof, x, y): [commitment - y G]_1 on the device
// (b200_check_proof_single_g1_batch), [s - x]_2 = SecretG2[1] - x GenG2 (:59-62) and the pairing check (:74) on the host.
extern "C" int b200_check_proof_single_batch(b200_ks* ks, const uint64_t* commitments, const uint64_t* proofs, const uint64_t* xs,
                                             const uint64_t* ys, size_t batch, uint8_t* ok) {
    G2J s2;
    CKS(ks_secret_g2(ks, 1, &s2));                                      // SecretG2[1]
    if (batch == 0) return B200_OK;
    std::vector<uint64_t> lhs(batch * 18);
    CKS(b200_check_proof_single_g1_batch(commitments, ys, batch, lhs.data()));
    return pairing_checks(lhs.data(), proofs, s2, xs, batch, ok);
}
extern "C" int b200_check_proof_single(b200_ks* ks, const uint64_t* commitment, const uint64_t* proof, const uint64_t* x, const uint64_t* y,
                                       int* ok) {
    uint8_t r = 0;
    *ok = 0;
    CKS(b200_check_proof_single_batch(ks, commitment, proof, x, y, 1, &r));
    *ok = r;
    return B200_OK;
}
// kzg_multi_proofs.go:47-88 CheckProofMulti for `batch` samples of n values: interpolation, MSM and subtraction on the device
// (b200_check_proof_multi_g1_batch), [s^n - x^n]_2 = SecretG2[n] - x^n GenG2 (:71-76) and the pairing check (:87) on the host.
extern "C" int b200_check_proof_multi_batch(b200_ks* ks, const uint64_t* commitments, const uint64_t* proofs, const uint64_t* xs,
                                            const uint64_t* ys, size_t n, size_t batch, uint8_t* ok) {
    G2J sn;
    CKS(ks_secret_g2(ks, n, &sn));                                      // SecretG2[len(ys)]
    if (batch == 0) return B200_OK;
    std::vector<uint64_t> lhs(batch * 18), xn(batch * 4);
    CKS(b200_check_proof_multi_g1_batch(ks, commitments, xs, ys, n, batch, lhs.data(), xn.data()));
    return pairing_checks(lhs.data(), proofs, sn, xn.data(), batch, ok);
}
extern "C" int b200_check_proof_multi(b200_ks* ks, const uint64_t* commitment, const uint64_t* proof, const uint64_t* x, const uint64_t* ys,
                                      size_t n, int* ok) {
    uint8_t r = 0;
    *ok = 0;
    CKS(b200_check_proof_multi_batch(ks, commitment, proof, x, ys, n, 1, &r));
    *ok = r;
    return B200_OK;
}

// Aggregated verification: `batch` proofs with ONE pairing check.  e(A_i, [1]_2) == e(proof_i, [t]_2 - c_i [1]_2) for all i
// (single: A_i = commitment_i - y_i G, t = s, c_i = x_i; multi: A_i = commitment_i - [I_i(s)]_1, t = s^n, c_i = x_i^n)
// is, for scalars r_i the prover cannot predict, equivalent (up to probability ~batch / r) to
//     e(sum r_i A_i + sum (r_i c_i) proof_i, [1]_2) == e(sum r_i proof_i, [t]_2):
// three MSMs of `batch` terms on the device (bucket MSM from 32 terms on) and one pairing on the host.  rs: batch canonical
// scalars from the caller's CSPRNG (the Go shim uses crypto/rand); a single r_i == 0 would drop proof i from the check,
// so zeros are rejected.
// a_pts, proofs, cs, rs: host arrays of `batch` entries; g_scalar (may be null): an extra term g_scalar * G on the left.
// Device-resident: both point arrays go up once, are checked to be on the curve there, and feed the three MSMs; three
// points come back.  The scalars r_i c_i are formed on the host (Fr products).
static int aggregate_check(const uint64_t* a_pts, const uint64_t* proofs, const uint64_t* cs, const uint64_t* rs, size_t batch,
                           const G2J& t2, const Fr* g_scalar_canon, int* ok) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    const size_t na = batch + (g_scalar_canon ? 1 : 0);
    std::vector<uint64_t> rc(batch * 4), ra(na * 4), apts(na * 18);
    for (size_t i = 0; i < batch; i++) {
        if (!fr_canon_valid(rs + 4 * i) || !fr_canon_valid(cs + 4 * i)) return B200_ERR_BAD_INPUT;
        const Fr r = fr_load_canon(rs + 4 * i);
        if (r.is_zero()) return B200_ERR_BAD_INPUT;
        fr_store_canon(rc.data() + 4 * i, fe_mul(fe_to_mont(r), fr_load_canon(cs + 4 * i)));
        if (!abi_coords_canonical(proofs + 18 * i, 3) || !abi_coords_canonical(a_pts + 18 * i, 3)) return B200_ERR_BAD_INPUT;
    }
    memcpy(ra.data(), rs, batch * 32);
    memcpy(apts.data(), a_pts, batch * 144);
    if (g_scalar_canon) {
        fr_store_canon(ra.data() + 4 * batch, *g_scalar_canon);
        g1_to_abi(apts.data() + 18 * batch, g1_generator());
    }
    CK(cudaSetDevice(g_device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf raw_a, raw_p, da, dp, k_a, k_rc, k_r, work, res, flag;
    CKS(raw_a.alloc(na * 144, st)); CKS(raw_p.alloc(batch * 144, st)); CKS(da.alloc(na * sizeof(G1J), st)); CKS(dp.alloc(batch * sizeof(G1J), st));
    CKS(k_a.alloc(na * 32, st)); CKS(k_rc.alloc(batch * 32, st)); CKS(k_r.alloc(batch * 32, st));
    CKS(work.alloc(3 * na * sizeof(G1J), st)); CKS(res.alloc(3 * 144, st)); CKS(flag.alloc(4, st));
    CK(cudaMemcpyAsync(raw_a.p, apts.data(), na * 144, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(raw_p.p, proofs, batch * 144, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(k_a.p, ra.data(), na * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(k_rc.p, rc.data(), batch * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(k_r.p, rs, batch * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(flag.p, 0, 4, st));
    launch_g1_from_abi(raw_a.as<uint64_t>(), da.as<G1J>(), na, st);
    launch_g1_from_abi(raw_p.as<uint64_t>(), dp.as<G1J>(), batch, st);
    launch_g1_on_curve(da.as<G1J>(), na, flag.as<uint32_t>(), st);
    launch_g1_on_curve(dp.as<G1J>(), batch, flag.as<uint32_t>(), st);
    G1J* w = work.as<G1J>();
    CKS(dev_lincomb(da.as<G1J>(), 0, k_a.as<Fr>(), 0, w, na, 1, st));                   // sum r_i A_i (+ g G)
    CKS(dev_lincomb(dp.as<G1J>(), 0, k_rc.as<Fr>(), 0, w + na, batch, 1, st));          // sum r_i c_i proof_i
    CKS(dev_lincomb(dp.as<G1J>(), 0, k_r.as<Fr>(), 0, w + 2 * na, batch, 1, st));       // sum r_i proof_i
    launch_g1_to_abi(w, res.as<uint64_t>(), 1, 1, 1, 1, 0, 0, st);
    launch_g1_to_abi(w + na, res.as<uint64_t>() + 18, 1, 1, 1, 1, 0, 0, st);
    launch_g1_to_abi(w + 2 * na, res.as<uint64_t>() + 36, 1, 1, 1, 1, 0, 0, st);
    CKS(check_launches());
    uint64_t h_res[54];
    uint32_t h_flag = 0;
    CK(cudaMemcpyAsync(h_res, res.p, sizeof h_res, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_flag, flag.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_flag) return B200_ERR_BAD_INPUT;                                              // a point off the curve
    const G1J lhs = g1_add(g1_from_abi(h_res), g1_from_abi(h_res + 18));
    *ok = pairings_verify(lhs, g2_generator(), g1_from_abi(h_res + 36), t2) ? 1 : 0;
    return B200_OK;
}
extern "C" int b200_check_proof_single_aggregate(b200_ks* ks, const uint64_t* commitments, const uint64_t* proofs, const uint64_t* xs,
                                                 const uint64_t* ys, const uint64_t* rs, size_t batch, int* ok) {
    *ok = 0;
    G2J s2;
    CKS(ks_secret_g2(ks, 1, &s2));
    if (batch == 0) { *ok = 1; return B200_OK; }
    // sum r_i (commitment_i - y_i G) = sum r_i commitment_i - (sum r_i y_i) G: one extra term instead of a product per proof
    Fr ry = Fr::zero();
    for (size_t i = 0; i < batch; i++) {
        if (!fr_canon_valid(ys + 4 * i) || !fr_canon_valid(rs + 4 * i)) return B200_ERR_BAD_INPUT;
        ry = fe_add(ry, fe_mul(fe_to_mont(fr_load_canon(rs + 4 * i)), fr_load_canon(ys + 4 * i)));
    }
    const Fr neg_ry = fe_neg(ry);
    return aggregate_check(commitments, proofs, xs, rs, batch, s2, &neg_ry, ok);
}
extern "C" int b200_check_proof_multi_aggregate(b200_ks* ks, const uint64_t* commitments, const uint64_t* proofs, const uint64_t* xs,
                                                const uint64_t* ys, size_t n, const uint64_t* rs, size_t batch, int* ok) {
    *ok = 0;
    G2J sn;
    CKS(ks_secret_g2(ks, n, &sn));
    if (batch == 0) { *ok = 1; return B200_OK; }
    std::vector<uint64_t> a(batch * 18), xn(batch * 4);
    CKS(b200_check_proof_multi_g1_batch(ks, commitments, xs, ys, n, batch, a.data(), xn.data()));
    return aggregate_check(a.data(), proofs, xn.data(), rs, batch, sn, nullptr, ok);
}

// ------------------------------------------------------------------------------ FK20
struct b200_fk {
    int device = 0;
    b200_ks* ks = nullptr;
    size_t n2 = 0, chunk_len = 1;
    G1J* d_x_ext_fft = nullptr;   // [chunk_len][n2 / chunk_len], natural order (kzg.go:62,110-114)
    size_t file_begin = 0, file_end = 1;   // chunk offsets ("files") held by this handle: all of them unless built sharded
    G1A* d_fb_table = nullptr;    // fixed-base window table over d_x_ext_fft (null when over budget)
    int fb_w = 8;                 // its window bits (8, or 4 for very large settings)
};

// kzg.go:43-64 / 73-116 + fk20_single.go:40-56 toeplitzPart1
// file_begin / file_end: the chunk offsets whose xExtFFT files (and window tables) this handle holds; a rank of the
// offset-sharded FK20 multi needs only its own (config 5 on 8 GPUs: 13 GB of tables per rank instead of 103 GB)
// x_ext_fft_host: previously exported files (b200_fk20_x_ext_fft) to adopt instead of recomputing them (setup cache)
static int fk20_settings_new(b200_ks* ks, size_t n2, size_t chunk_len, b200_fk** out, size_t file_begin = 0, size_t file_end = (size_t)-1,
                             const uint64_t* x_ext_fft_host = nullptr) {
    *out = nullptr;
    b200_fs* fs = ks->fs;
    if (n2 > fs->max_width) return B200_ERR_TOO_LARGE;       // kzg.go:44-46 / 74-76
    if (!is_pow2(n2)) return B200_ERR_NOT_POW2;              // kzg.go:47-49 / 77-79
    if (n2 < 2) return B200_ERR_TOO_SMALL;                   // kzg.go:50-52 / 80-82
    if (chunk_len > n2 / 2) return B200_ERR_TOO_LARGE;       // kzg.go:83-85
    if (!is_pow2(chunk_len)) return B200_ERR_NOT_POW2;       // kzg.go:86-91
    if (n2 / chunk_len > ((size_t)1 << 24)) return B200_ERR_TOO_LARGE;   // Fr NTT limit of this library (launch_fr_ntt)
    CK(cudaSetDevice(fs->device));
    const size_t n = n2 / 2, l = chunk_len, k = n / l, k2 = 2 * k;
    const unsigned logk2 = log2u(k2);
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    b200_fk* fk = new (std::nothrow) b200_fk();
    if (!fk) return B200_ERR_CUDA;
    fk->ks = ks; fk->n2 = n2; fk->chunk_len = l; fk->device = fs->device;
    if (file_end == (size_t)-1) file_end = l;
    fk->file_begin = file_begin; fk->file_end = file_end;
    const size_t m = file_end - file_begin;            // files held
    int rc = B200_OK;
    do {
        if (cudaMalloc(&fk->d_x_ext_fft, (m ? m : 1) * k2 * sizeof(G1J)) != cudaSuccess) { rc = B200_ERR_CUDA; break; }
        if (m == 0) break;
        DevBuf work;
        if ((rc = work.alloc(m * k2 * sizeof(G1J), st))) break;
        if (x_ext_fft_host) {
            if (cudaMemcpyAsync(work.p, x_ext_fft_host, m * k2 * 144, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = B200_ERR_CUDA; break; }
            launch_g1_from_abi(work.as<uint64_t>(), fk->d_x_ext_fft, m * k2, st);
            if ((rc = check_launches())) break;
            if (cudaStreamSynchronize(st) != cudaSuccess) { g_cuda_err = "sync in fk20 settings"; rc = B200_ERR_CUDA; break; }
            if ((rc = build_fixed_base(fk->d_x_ext_fft, m * k2, &fk->d_fb_table, &fk->fb_w, st))) break;
            break;
        }
        launch_g1_fill_infinity(work.as<G1J>(), m * k2, st);
        // file `off`: x[i] = SecretG1[n - l - 1 - off - i l], i < k - 1; x[k-1 .. 2k-1] = infinity
        launch_fk20_gather_x(ks->d_secret_g1, work.as<G1J>(), n, l, file_begin, m, st);
        // FFT_G1 of every file (forward), natural order result
        if ((rc = dev_g1_fft_stages(fs, work.as<G1J>(), logk2, m, 1, k2, false, true, st))) break;
        launch_g1_copy(fk->d_x_ext_fft, 1, k2, work.as<G1J>(), 1, k2, k2, m, 1, logk2, st);
        if ((rc = check_launches())) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) { g_cuda_err = "sync in fk20 settings"; rc = B200_ERR_CUDA; break; }
        if ((rc = build_fixed_base(fk->d_x_ext_fft, m * k2, &fk->d_fb_table, &fk->fb_w, st))) break;
    } while (0);
    if (rc) { b200_fk20_settings_free(fk); return rc; }
    *out = fk;
    return B200_OK;
}
extern "C" int b200_fk20_single_settings_new(b200_ks* ks, size_t n2, b200_fk** out) { return fk20_settings_new(ks, n2, 1, out); }
extern "C" int b200_fk20_multi_settings_new(b200_ks* ks, size_t n2, size_t chunk_len, b200_fk** out) {
    if (chunk_len < 1) { *out = nullptr; return B200_ERR_TOO_SMALL; }   // kzg.go:89-91
    return fk20_settings_new(ks, n2, chunk_len, out);
}
// kzg.go:73-116 for the ranks of an offset-sharded FK20 multi: only the files [off_begin, off_end) are built and kept
extern "C" int b200_fk20_multi_settings_new_sharded(b200_ks* ks, size_t n2, size_t chunk_len, size_t off_begin, size_t off_end, b200_fk** out) {
    *out = nullptr;
    if (chunk_len < 1) return B200_ERR_TOO_SMALL;
    if (off_begin > off_end || off_end > chunk_len) return B200_ERR_BAD_INPUT;
    return fk20_settings_new(ks, n2, chunk_len, out, off_begin, off_end);
}
// Settings from cached xExtFFT files (the output of b200_fk20_x_ext_fft for the offsets [off_begin, off_end), concatenated):
// skips the G1 transforms of kzg.go:57-62 / 101-114; the window tables are rebuilt (seconds of device time, faster than reading them).
// The caller vouches that the files belong to `ks` (key the cache by a digest of the setup, as go_kzg_b200/kzg.py does).
extern "C" int b200_fk20_settings_new_from_x_ext_fft(b200_ks* ks, size_t n2, size_t chunk_len, size_t off_begin, size_t off_end,
                                                     const uint64_t* x_ext_fft, b200_fk** out) {
    *out = nullptr;
    if (chunk_len < 1) return B200_ERR_TOO_SMALL;
    if (off_begin > off_end || off_end > chunk_len || !x_ext_fft) return B200_ERR_BAD_INPUT;
    return fk20_settings_new(ks, n2, chunk_len, out, off_begin, off_end, x_ext_fft);
}
extern "C" void b200_fk20_settings_free(b200_fk* fk) {
    if (!fk) return;
    cudaSetDevice(fk->device);
    cudaFree(fk->d_x_ext_fft);
    cudaFree(fk->d_fb_table);
    delete fk;
}
extern "C" int b200_fk20_x_ext_fft(b200_fk* fk, size_t file, uint64_t* out) {
    if (file < fk->file_begin || file >= fk->file_end) return B200_ERR_BAD_INPUT;
    CK(cudaSetDevice(fk->ks->fs->device));
    const size_t k2 = fk->n2 / fk->chunk_len;
    file -= fk->file_begin;
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf raw;
    CKS(raw.alloc(k2 * 144, st));
    launch_g1_to_abi(fk->d_x_ext_fft + file * k2, raw.as<uint64_t>(), k2, 1, 1, k2, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out, raw.p, k2 * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
// kernels launched by the last FK20 / commit+FK20 call made on the calling thread (handles stay immutable)
static thread_local unsigned long long t_last_launches = 0;
extern "C" uint64_t b200_last_launch_count(void) { return t_last_launches; }
extern "C" uint64_t b200_fk20_last_launch_count(const b200_fk*) { return t_last_launches; }

// The FK20 pipeline on device buffers, for `batch` polynomials of n coefficients (canonical):
//   c      = toeplitz coefficients (per chunk offset)                      fk20_single.go:89-119
//   c^     = NTT(c) / 2k            (the 1/2k of the later inverse G1 transform folded in here)
//   hExt^  = sum over offsets of c^ (.) xExtFFT[offset]                     fk20_single.go:59-77, fk20_multi.go:80-91
//   h      = IFFT_G1(hExt^)[:k]     decimation in frequency: natural in, bit-reversed out, so
//                                   h[m] (m < k) lands on even slot 2 rev_k(m)
//   proofs = FFT_G1(h)              (mode 0: k points, the even slots *are* the bit-reversed input of a DIT)
//          | FFT_G1(h ++ 0^k)       (mode 1/2: 2k points; odd slots are cleared first)
// mode 0: FK20Single (natural order); 1: *DAOptimized (natural order, 2k proofs); 2: DAUsing* (reverse bit order)
// d_comp != nullptr: the proofs leave as 48-byte compressed points (d_proofs unused)
static int dev_fk20(b200_fk* fk, const uint64_t* d_polys, size_t n, size_t batch, int mode, uint64_t* d_proofs, cudaStream_t st,
                    uint8_t* d_comp = nullptr) {
    b200_fs* fs = fk->ks->fs;
    const size_t l = fk->chunk_len, k = n / l, k2 = 2 * k;
    if (fk->file_begin != 0 || fk->file_end != l) return B200_ERR_BAD_INPUT;   // sharded settings hold some of the files only
    const unsigned logk2 = log2u(k2);
    DevBuf c, h, tmp;
    CKS(c.alloc(batch * l * k2 * sizeof(Fr), st));
    CKS(h.alloc(batch * l * k2 * sizeof(G1J), st));
    if (l == 1) launch_toeplitz_coeffs(d_polys, c.as<Fr>(), n, batch, st);
    else launch_toeplitz_coeffs_strided(d_polys, c.as<Fr>(), n, l, batch, st);
    if (logk2 > 12) CKS(tmp.alloc(batch * l * k2 * sizeof(Fr), st));
    Fr scale = fr_inv_of_u64(k2);
    launch_fr_ntt(fs->dom, c.as<Fr>(), c.as<Fr>(), tmp.as<Fr>(), logk2, batch * l, false, &scale, st);
    if (mode == 0) {
        // even entries carry 1/2 instead of 1/2k (see below): times k
        launch_fr_mul_even_odd(c.as<Fr>(), fr_from_u64(k), Fr::one(), batch * l * k2, st);
    }
    // FK20Single with a window table: ToeplitzPart2 and the first two inverse stages over the odd slots in one kernel
    const bool fold2 = mode == 0 && l == 1 && fk->d_fb_table && k >= 8;
    if (fold2) launch_fk20_part2_fold2(fk->d_fb_table, fk->fb_w, c.as<Fr>(), fs->dom.reverse, fs->max_width / k, h.as<G1J>(), k2, k, batch, st);
    else if (fk->d_fb_table) launch_g1_mul_fixed_base(fk->d_fb_table, fk->fb_w, c.as<Fr>(), 1, h.as<G1J>(), l * k2, l * k2, batch, st);
    else launch_g1_mul_var(fk->d_x_ext_fft, 0, c.as<Fr>(), 1, h.as<G1J>(), l * k2, l * k2, batch, st);
    for (size_t cnt = l; cnt > 1; cnt /= 2) launch_g1_fold(h.as<G1J>(), l * k2, (cnt / 2) * k2, cnt * k2, batch, st);
    CKS(check_launches());
    const size_t bstride = l * k2;
    if (mode == 0) {
        // FK20Single: only FFT_k(h[:k]) is wanted.  With H = hExt^ (2k evaluations of a polynomial
        // h_lo + X^k h_hi) the even entries evaluate h_lo + h_hi on the order-k domain and the odd
        // ones evaluate (h_lo - h_hi)(g X), g = w_2k.  Hence
        //     FFT_k(h_lo) = 1/2 H_even + 1/2 FFT_k( g^-m . IFFT_k(H_odd)[m] ),
        // two size-k transforms and k twists instead of a size-2k and a size-k transform
        // (about 0.69x the twiddle multiplications).  1/2 and 1/(2k) were folded into c^ above.
        // The odd slots (stride 2) are transformed in place: DIF inverse (bit-reversed result),
        // twist by position, DIT forward; then the even slots are added.
        const unsigned logk = logk2 - 1;
        G1J* odd = h.as<G1J>() + 1;
        CKS(dev_g1_fft_stages(fs, odd, logk, batch, 2, bstride, true, true, st, fold2 ? 2 : 0));
        const ScalarProgram* inv_progs;
        CKS(fs_programs(fs, 1, program_mode_for_batch(batch), &inv_progs));
        launch_g1_mul_programs(odd, k, batch, 2, bstride, inv_progs, (fs->max_width / 2) / k, 1, logk, st);
        CKS(dev_g1_fft_stages(fs, odd, logk, batch, 2, bstride, false, false, st));
        launch_g1_add_arrays(odd, 2, bstride, h.as<G1J>(), 2, bstride, k, batch, st);
        if (d_comp) launch_g1_compress(odd, d_comp, k, batch, 2, bstride, 0, 0, st);
        else launch_g1_to_abi(odd, d_proofs, k, batch, 2, bstride, 0, 0, st);
    } else {
        CKS(dev_g1_fft_stages(fs, h.as<G1J>(), logk2, batch, 1, bstride, true, true, st));
        // clear the odd slots (the discarded upper half of the inverse transform): h ++ 0^k
        G1J* d_inf = nullptr;
        CKS(dev_infinity(&d_inf));
        launch_g1_copy(h.as<G1J>() + 1, 2, bstride, d_inf, 0, 0, k, batch, 0, 0, st);
        CKS(dev_g1_fft_stages(fs, h.as<G1J>(), logk2, batch, 1, bstride, false, false, st));
        if (d_comp) launch_g1_compress(h.as<G1J>(), d_comp, k2, batch, 1, bstride, mode == 2 ? 1 : 0, logk2, st);
        else launch_g1_to_abi(h.as<G1J>(), d_proofs, k2, batch, 1, bstride, mode == 2 ? 1 : 0, logk2, st);
    }
    return check_launches();
}

// host-buffer front end shared by the five FK20 entry points
static int host_fk20(b200_fk* fk, const uint64_t* poly, size_t n, int mode, uint64_t* proofs) {
    CK(cudaSetDevice(fk->ks->fs->device));
    const size_t n_out = mode == 0 ? n / fk->chunk_len : 2 * n / fk->chunk_len;
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf dp, dout;
    CKS(dp.alloc(n * 32, st)); CKS(dout.alloc(n_out * 144, st));
    CK(cudaMemcpyAsync(dp.p, poly, n * 32, cudaMemcpyHostToDevice, st));
    unsigned long long before = g_launch_count;
    CKS(dev_fk20(fk, dp.as<uint64_t>(), n, 1, mode, dout.as<uint64_t>(), st));
    t_last_launches = g_launch_count - before;
    CK(cudaMemcpyAsync(proofs, dout.p, n_out * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
static bool upper_half_zero(const uint64_t* poly, size_t n2) {
    for (size_t i = (n2 / 2) * 4; i < n2 * 4; i++) if (poly[i]) return false;
    return true;
}

extern "C" int b200_fk20_single(b200_fk* fk, const uint64_t* poly, size_t n, uint64_t* proofs) {
    if (fk->chunk_len != 1) return B200_ERR_BAD_INPUT;
    if (2 * n != fk->n2) return B200_ERR_LEN_MISMATCH;   // fk20_single.go:60-62 (toeplitz coeffs vs xExtFFT)
    return host_fk20(fk, poly, n, 0, proofs);
}
extern "C" int b200_fk20_single_da_optimized(b200_fk* fk, const uint64_t* poly, size_t n2, uint64_t* proofs) {
    if (fk->chunk_len != 1) return B200_ERR_BAD_INPUT;
    if (n2 > fk->ks->fs->max_width) return B200_ERR_TOO_LARGE;   // fk20_single.go:140-145
    if (!is_pow2(n2)) return B200_ERR_NOT_POW2;                  // fk20_single.go:146-149
    if (!upper_half_zero(poly, n2)) return B200_ERR_BAD_INPUT;   // fk20_single.go:150-154
    if (n2 != fk->n2) return B200_ERR_LEN_MISMATCH;
    return host_fk20(fk, poly, n2 / 2, 1, proofs);
}
extern "C" int b200_da_using_fk20(b200_fk* fk, const uint64_t* poly, size_t n, uint64_t* proofs) {
    if (fk->chunk_len != 1) return B200_ERR_BAD_INPUT;
    if (n > fk->ks->fs->max_width / 2) return B200_ERR_TOO_LARGE;   // fk20_single.go:178-180
    if (!is_pow2(n)) return B200_ERR_NOT_POW2;                      // fk20_single.go:181-183
    if (2 * n != fk->n2) return B200_ERR_LEN_MISMATCH;
    return host_fk20(fk, poly, n, 2, proofs);
}
// DAUsingFK20 for `batch` polynomials (fk20_single.go:176-196 per polynomial): lanes of a warp = polynomials, like the
// headline batch call.  Device-buffer form on the caller's stream, and the host-buffer form around it.
extern "C" int b200_da_using_fk20_batch_dev(b200_fk* fk, const void* d_polys, size_t n, size_t batch, void* d_proofs, void* cuda_stream) {
    if (fk->chunk_len != 1) return B200_ERR_BAD_INPUT;
    if (n > fk->ks->fs->max_width / 2) return B200_ERR_TOO_LARGE;   // fk20_single.go:178-180
    if (!is_pow2(n)) return B200_ERR_NOT_POW2;                      // fk20_single.go:181-183
    if (2 * n != fk->n2) return B200_ERR_LEN_MISMATCH;
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fk->device));
    unsigned long long before = g_launch_count;
    CKS(dev_fk20(fk, (const uint64_t*)d_polys, n, batch, 2, (uint64_t*)d_proofs, (cudaStream_t)cuda_stream));
    t_last_launches = g_launch_count - before;
    return B200_OK;
}
extern "C" int b200_da_using_fk20_batch(b200_fk* fk, const uint64_t* polys, size_t n, size_t batch, uint64_t* proofs) {
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fk->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf dp, dout;
    CKS(dp.alloc(batch * n * 32, st)); CKS(dout.alloc(batch * 2 * n * 144, st));
    CK(cudaMemcpyAsync(dp.p, polys, batch * n * 32, cudaMemcpyHostToDevice, st));
    CKS(b200_da_using_fk20_batch_dev(fk, dp.p, n, batch, dout.p, st));
    CK(cudaMemcpyAsync(proofs, dout.p, batch * 2 * n * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
extern "C" int b200_fk20_multi_da_optimized(b200_fk* fk, const uint64_t* poly, size_t n2, uint64_t* proofs) {
    if (n2 > fk->ks->fs->max_width) return B200_ERR_TOO_LARGE;   // fk20_multi.go:60-63
    if (!upper_half_zero(poly, n2)) return B200_ERR_BAD_INPUT;   // fk20_multi.go:65-69
    if (n2 != fk->n2) return B200_ERR_LEN_MISMATCH;
    return host_fk20(fk, poly, n2 / 2, 1, proofs);
}
extern "C" int b200_da_using_fk20_multi(b200_fk* fk, const uint64_t* poly, size_t n, uint64_t* proofs) {
    if (n > fk->ks->fs->max_width / 2) return B200_ERR_TOO_LARGE;   // fk20_multi.go:115-117
    if (!is_pow2(n)) return B200_ERR_NOT_POW2;                      // fk20_multi.go:118-120
    if (2 * n != fk->n2) return B200_ERR_LEN_MISMATCH;
    return host_fk20(fk, poly, n, 2, proofs);
}

// ------------------------------------------------------------------------------ multi-GPU building blocks
extern "C" int b200_fk20_multi_partial_dev(b200_fk* fk, const void* d_poly, size_t n, size_t off_begin, size_t off_end,
                                           void* d_partial, void* cuda_stream) {
    if (2 * n != fk->n2) return B200_ERR_LEN_MISMATCH;
    const size_t l = fk->chunk_len, k = n / l, k2 = 2 * k;
    if (off_begin > off_end || off_end > l) return B200_ERR_BAD_INPUT;
    if (off_begin < off_end && (off_begin < fk->file_begin || off_end > fk->file_end)) return B200_ERR_BAD_INPUT;   // files this handle does not hold
    CK(cudaSetDevice(fk->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    b200_fs* fs = fk->ks->fs;
    const size_t m = off_end - off_begin;
    if (m == 0) { launch_g1_fill_infinity((G1J*)d_partial, k2, st); return check_launches(); }   // all-zero bytes == infinity in both encodings
    const unsigned logk2 = log2u(k2);
    DevBuf c, h, tmp;
    CKS(c.alloc(l * k2 * sizeof(Fr), st));
    CKS(h.alloc(m * k2 * sizeof(G1J), st));
    launch_toeplitz_coeffs_strided((const uint64_t*)d_poly, c.as<Fr>(), n, l, 1, st);   // every offset; only [off_begin, off_end) is used
    if (logk2 > 12) CKS(tmp.alloc(m * k2 * sizeof(Fr), st));
    Fr scale = fr_inv_of_u64(k2);     // the inverse transform's 1/2k, as in dev_fk20
    Fr* c_mine = c.as<Fr>() + off_begin * k2;
    launch_fr_ntt(fs->dom, c_mine, c_mine, tmp.as<Fr>(), logk2, m, false, &scale, st);
    const size_t rel = off_begin - fk->file_begin;
    if (fk->d_fb_table) launch_g1_mul_fixed_base(fk->d_fb_table + rel * k2 * fixed_base_row_entries(fk->fb_w), fk->fb_w, c_mine, 1, h.as<G1J>(), m * k2, m * k2, 1, st);
    else launch_g1_mul_var(fk->d_x_ext_fft + rel * k2, 0, c_mine, 1, h.as<G1J>(), m * k2, m * k2, 1, st);
    // sum the m files: fold the tail onto the head until one file is left (m need not be a power of two)
    for (size_t cnt = m; cnt > 1;) {
        size_t half = (cnt + 1) / 2;
        launch_g1_fold(h.as<G1J>(), m * k2, half * k2, cnt * k2, 1, st);
        cnt = half;
    }
    launch_g1_to_abi(h.as<G1J>(), (uint64_t*)d_partial, k2, 1, 1, k2, 0, 0, st);
    return check_launches();
}

extern "C" int b200_g1_sum_dev(const void* d_parts, size_t n_parts, size_t count, void* d_out, void* cuda_stream) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (n_parts == 0) { launch_g1_fill_infinity((G1J*)d_out, count, st); return check_launches(); }
    DevBuf w;
    CKS(w.alloc(n_parts * count * sizeof(G1J), st));
    launch_g1_from_abi((const uint64_t*)d_parts, w.as<G1J>(), n_parts * count, st);
    for (size_t cnt = n_parts; cnt > 1;) {
        size_t half = (cnt + 1) / 2;
        launch_g1_fold(w.as<G1J>(), n_parts * count, half * count, cnt * count, 1, st);
        cnt = half;
    }
    launch_g1_to_abi(w.as<G1J>(), (uint64_t*)d_out, count, 1, 1, count, 0, 0, st);
    return check_launches();
}

extern "C" int b200_fk20_multi_finish_dev(b200_fk* fk, const void* d_h_ext_fft, int reverse_bits, void* d_proofs, void* cuda_stream) {
    CK(cudaSetDevice(fk->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    b200_fs* fs = fk->ks->fs;
    const size_t k2 = fk->n2 / fk->chunk_len, k = k2 / 2;
    const unsigned logk2 = log2u(k2);
    DevBuf h;
    CKS(h.alloc(k2 * sizeof(G1J), st));
    launch_g1_from_abi((const uint64_t*)d_h_ext_fft, h.as<G1J>(), k2, st);
    CKS(dev_g1_fft_stages(fs, h.as<G1J>(), logk2, 1, 1, k2, true, true, st));     // fk20_multi.go:93 (1/2k already folded in)
    G1J* d_inf = nullptr;
    CKS(dev_infinity(&d_inf));
    launch_g1_copy(h.as<G1J>() + 1, 2, k2, d_inf, 0, 0, k, 1, 0, 0, st);          // fk20_multi.go:100-103
    CKS(dev_g1_fft_stages(fs, h.as<G1J>(), logk2, 1, 1, k2, false, false, st));   // fk20_multi.go:104
    launch_g1_to_abi(h.as<G1J>(), (uint64_t*)d_proofs, k2, 1, 1, k2, reverse_bits ? 1 : 0, logk2, st);
    return check_launches();
}

// Sharded form of b200_fk20_multi_finish_dev for `world` = 2^s ranks that all hold the summed hExtFFT
// (fk20_multi.go:93-106 on several GPUs).  A DIF transform splits after s stages into 2^s independent
// transforms on contiguous blocks, a DIT transform (bit-reversed input) is block-local until its last s stages:
//   local:  rank r takes the input down to its block with s half stages (only the outputs it needs), runs the
//           remaining inverse stages, clears the odd slots (= the upper half of h in natural order,
//           fk20_multi.go:100-103) and runs the block-local forward stages;  d_block: k2 / world internal points
//   merge:  after an all-gather of the blocks in rank order, the last s forward stages and the output conversion.
// Per rank ~ (1 + 2 (log2 k2 - s) / 2^s + 3 s) k2 / 2 twiddle products instead of 2 log2 k2 * k2 / 2.
extern "C" int b200_fk20_multi_finish_local_dev(b200_fk* fk, const void* d_h_ext_fft, size_t rank, size_t world, void* d_block,
                                                void* cuda_stream) {
    CK(cudaSetDevice(fk->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    b200_fs* fs = fk->ks->fs;
    const size_t k2 = fk->n2 / fk->chunk_len;
    const unsigned logk2 = log2u(k2);
    if (!is_pow2(world) || rank >= world || world * 2 > k2) return B200_ERR_BAD_INPUT;
    const unsigned s = log2u(world);
    const size_t blk = k2 >> s;
    const unsigned logblk = logk2 - s;
    const ScalarProgram *inv_progs, *fwd_progs;
    CKS(fs_programs(fs, 1, 0, &inv_progs));
    CKS(fs_programs(fs, 0, 0, &fwd_progs));
    const size_t halfw = fs->max_width / 2;
    DevBuf h, t0, t1;
    CKS(h.alloc(k2 * sizeof(G1J), st));
    launch_g1_from_abi((const uint64_t*)d_h_ext_fft, h.as<G1J>(), k2, st);
    G1J* cur = h.as<G1J>();
    if (s > 0) { CKS(t0.alloc((k2 / 2) * sizeof(G1J), st)); CKS(t1.alloc((k2 / 4 ? k2 / 4 : 1) * sizeof(G1J), st)); }
    size_t m = k2 / 2;
    for (unsigned lvl = 0; lvl < s; lvl++, m >>= 1) {
        const int lower = (int)((rank >> (s - 1 - lvl)) & 1);             // block index, most significant bit first
        G1J* out = (lvl + 1 == s) ? (G1J*)d_block : ((lvl & 1) ? t1.as<G1J>() : t0.as<G1J>());
        launch_g1_dif_half_stage(cur, out, m, lower, inv_progs, halfw / m, st);
        cur = out;
    }
    G1J* block = (G1J*)d_block;
    if (s == 0) CK(cudaMemcpyAsync(block, cur, k2 * sizeof(G1J), cudaMemcpyDeviceToDevice, st));
    StagePrograms spi, spf;
    CKS(fs_stage_programs(fs, 1, 1, &spi));
    CKS(fs_stage_programs(fs, 0, 1, &spf));
    for (size_t mm = blk / 2; mm >= 1; mm >>= 1) launch_stage_auto(spi, block, blk / 2, 1, mm, 1, blk, true, halfw / mm, st);
    G1J* d_inf = nullptr;
    CKS(dev_infinity(&d_inf));
    launch_g1_copy(block + 1, 2, blk, d_inf, 0, 0, blk / 2, 1, 0, 0, st);
    for (size_t mm = 1; mm <= blk / 2; mm <<= 1) launch_stage_auto(spf, block, blk / 2, 1, mm, 1, blk, false, halfw / mm, st);
    (void)logblk;
    return check_launches();
}
extern "C" int b200_fk20_multi_finish_merge_dev(b200_fk* fk, const void* d_blocks, size_t world, int reverse_bits, void* d_proofs,
                                                void* cuda_stream) {
    CK(cudaSetDevice(fk->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    b200_fs* fs = fk->ks->fs;
    const size_t k2 = fk->n2 / fk->chunk_len;
    const unsigned logk2 = log2u(k2);
    if (!is_pow2(world) || world * 2 > k2) return B200_ERR_BAD_INPUT;
    const ScalarProgram* fwd_progs;
    CKS(fs_programs(fs, 0, 0, &fwd_progs));
    const size_t halfw = fs->max_width / 2;
    DevBuf h;
    CKS(h.alloc(k2 * sizeof(G1J), st));
    CK(cudaMemcpyAsync(h.p, d_blocks, k2 * sizeof(G1J), cudaMemcpyDeviceToDevice, st));
    StagePrograms spf;
    CKS(fs_stage_programs(fs, 0, 1, &spf));
    for (size_t mm = k2 / world; mm <= k2 / 2; mm <<= 1) launch_stage_auto(spf, h.as<G1J>(), k2 / 2, 1, mm, 1, k2, false, halfw / mm, st);
    launch_g1_to_abi(h.as<G1J>(), (uint64_t*)d_proofs, k2, 1, 1, k2, reverse_bits ? 1 : 0, logk2, st);
    return check_launches();
}

// The merge, sharded as well: after the all-gather of the blocks every rank holds the whole array, and the last s forward
// stages only ever combine the `world` elements i, i + blk, i + 2 blk, .. of one position i inside the blocks.  Rank r takes
// the positions [r sub, (r + 1) sub), sub = blk / world, of every block: `sub` independent size-`world` transforms with
// element stride blk whose twiddles depend on the position (index (c_low blk + i) of the stage's table) -- 1 / world of the
// butterflies per rank instead of all of them.  d_blocks is overwritten; d_part receives [world][sub] internal points.
// A second all-gather of the parts (rank-major) and b200_fk20_multi_finish_assemble_dev give every rank the full result.
extern "C" int b200_fk20_multi_finish_merge_part_dev(b200_fk* fk, void* d_blocks, size_t rank, size_t world, void* d_part, void* cuda_stream) {
    CK(cudaSetDevice(fk->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    b200_fs* fs = fk->ks->fs;
    const size_t k2 = fk->n2 / fk->chunk_len;
    if (!is_pow2(world) || rank >= world || world * world > k2) return B200_ERR_BAD_INPUT;
    const size_t blk = k2 / world, sub = blk / world, halfw = fs->max_width / 2;
    const ScalarProgram* fwd_progs;
    CKS(fs_programs(fs, 0, 0, &fwd_progs));                    // lanes hold different positions: per-lane fixed-window programs
    G1J* mine = (G1J*)d_blocks + rank * sub;
    for (size_t t = 1; t < world; t <<= 1)                     // stage with half length mm = blk t pairs block c with block c + t
        launch_g1_fft_stage(mine, world / 2, sub, t, blk, 1, false, fwd_progs, halfw / (blk * t), st, 0, blk, 1, rank * sub);
    launch_g1_copy((G1J*)d_part, 1, sub, mine, 1, blk, sub, world, 0, 0, st);   // part[c][i] = blocks[c blk + r sub + i]
    return check_launches();
}
extern "C" int b200_fk20_multi_finish_assemble_dev(b200_fk* fk, const void* d_parts, size_t world, int reverse_bits, void* d_proofs,
                                                   void* cuda_stream) {
    CK(cudaSetDevice(fk->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t k2 = fk->n2 / fk->chunk_len;
    if (!is_pow2(world) || world * world > k2) return B200_ERR_BAD_INPUT;
    const size_t sub = k2 / world / world;
    DevBuf h;
    CKS(h.alloc(k2 * sizeof(G1J), st));
    // gathered parts are [r][c][i]; natural position is (c, r, i)
    launch_g1_swap_digits(h.as<G1J>(), (const G1J*)d_parts, world, world, sub, st);
    launch_g1_to_abi(h.as<G1J>(), (uint64_t*)d_proofs, k2, 1, 1, k2, reverse_bits ? 1 : 0, log2u(k2), st);
    return check_launches();
}

extern "C" int b200_commit_partial_dev(b200_ks* ks, const void* d_coeffs, size_t begin, size_t end, void* d_out, void* cuda_stream) {
    if (begin > end || end > ks->n_g1) return B200_ERR_BAD_INPUT;
    CK(cudaSetDevice(ks->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t m = end - begin;
    if (m == 0) { launch_g1_fill_infinity((G1J*)d_out, 1, st); return check_launches(); }
    DevBuf work;
    CKS(work.alloc(m * sizeof(G1J), st));
    CKS(dev_lincomb(ks->d_secret_g1 + begin, 0, (const Fr*)d_coeffs + begin, 0, work.as<G1J>(), m, 1, st));
    launch_g1_to_abi(work.as<G1J>(), (uint64_t*)d_out, 1, 1, 1, m, 0, 0, st);
    return check_launches();
}

extern "C" int b200_generate_testing_setup_g1(const uint64_t* secret, size_t n, uint64_t* out) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    if (n == 0) return B200_OK;
    CK(cudaSetDevice(g_device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    Fr sq[40];
    sq[0] = fr_from_abi_mont(secret);
    for (int j = 1; j < 40; j++) sq[j] = fe_mul(sq[j - 1], sq[j - 1]);
    G1J gen = g1_generator();
    DevBuf dsq, dk, dg, dpts, raw;
    CKS(dsq.alloc(sizeof sq, st)); CKS(dk.alloc(n * sizeof(Fr), st)); CKS(dg.alloc(sizeof(G1J), st));
    CKS(dpts.alloc(n * sizeof(G1J), st)); CKS(raw.alloc(n * 144, st));
    CK(cudaMemcpyAsync(dsq.p, sq, sizeof sq, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dg.p, &gen, sizeof gen, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    launch_fr_powers(dsq.as<Fr>(), dk.as<Fr>(), n, st);
    // one shared base, n scalars: (n = 1, batch = n) with a zero base stride
    launch_g1_mul_var(dg.as<G1J>(), 0, dk.as<Fr>(), 0, dpts.as<G1J>(), 1, 1, n, st);
    launch_g1_to_abi(dpts.as<G1J>(), raw.as<uint64_t>(), n, 1, 1, n, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out, raw.p, n * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// ------------------------------------------------------------------------------ headline unit
extern "C" int b200_commit_fk20_batch_dev(b200_fk* fk, const void* d_polys, size_t n, size_t batch, void* d_commitments,
                                          void* d_proofs, void* cuda_stream) {
    if (fk->chunk_len != 1) return B200_ERR_BAD_INPUT;
    if (2 * n != fk->n2) return B200_ERR_LEN_MISMATCH;
    if (n > fk->ks->n_g1) return B200_ERR_LEN_MISMATCH;
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fk->ks->fs->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    unsigned long long before = g_launch_count;
    {
        DevBuf work;
        CKS(work.alloc(batch * n * sizeof(G1J), st));
        const G1A* fb = nullptr;
        int fbw = 8;
        CKS(ks_fixed_base(fk->ks, n, &fb, &fbw, st));
        CKS(dev_lincomb(fk->ks->d_secret_g1, 0, (const Fr*)d_polys, 0, work.as<G1J>(), n, batch, st, fb, fbw));
        launch_g1_to_abi(work.as<G1J>(), (uint64_t*)d_commitments, 1, batch, 1, n, 0, 0, st);
    }
    CKS(dev_fk20(fk, (const uint64_t*)d_polys, n, batch, 0, (uint64_t*)d_proofs, st));
    t_last_launches = g_launch_count - before;
    return B200_OK;
}
extern "C" int b200_commit_fk20_batch(b200_fk* fk, const uint64_t* polys, size_t n, size_t batch, uint64_t* commitments,
                                      uint64_t* proofs) {
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fk->ks->fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf dp, dc, dout;
    CKS(dp.alloc(batch * n * 32, st)); CKS(dc.alloc(batch * 144, st)); CKS(dout.alloc(batch * n * 144, st));
    CK(cudaMemcpyAsync(dp.p, polys, batch * n * 32, cudaMemcpyHostToDevice, st));
    CKS(b200_commit_fk20_batch_dev(fk, dp.p, n, batch, dc.p, dout.p, st));
    CK(cudaMemcpyAsync(commitments, dc.p, batch * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proofs, dout.p, batch * n * 144, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}

// ---- compressed outputs and the eth blob path (rank 1) -----------------------
static int commit_fk20_compressed_dev(b200_fk* fk, const void* d_polys, size_t n, size_t batch, uint8_t* d_commit48,
                                      uint8_t* d_proofs48, cudaStream_t st) {
    if (fk->chunk_len != 1) return B200_ERR_BAD_INPUT;
    if (2 * n != fk->n2) return B200_ERR_LEN_MISMATCH;
    if (n > fk->ks->n_g1) return B200_ERR_LEN_MISMATCH;
    if (batch == 0) return B200_OK;
    unsigned long long before = g_launch_count;
    {
        DevBuf work;
        CKS(work.alloc(batch * n * sizeof(G1J), st));
        const G1A* fb = nullptr;
        int fbw = 8;
        CKS(ks_fixed_base(fk->ks, n, &fb, &fbw, st));
        CKS(dev_lincomb(fk->ks->d_secret_g1, 0, (const Fr*)d_polys, 0, work.as<G1J>(), n, batch, st, fb, fbw));
        launch_g1_compress(work.as<G1J>(), d_commit48, 1, batch, 1, n, 0, 0, st);
    }
    CKS(dev_fk20(fk, (const uint64_t*)d_polys, n, batch, 0, nullptr, st, d_proofs48));
    t_last_launches = g_launch_count - before;
    return B200_OK;
}
extern "C" int b200_commit_fk20_batch_compressed_dev(b200_fk* fk, const void* d_polys, size_t n, size_t batch, void* d_commitments48,
                                                     void* d_proofs48, void* cuda_stream) {
    CK(cudaSetDevice(fk->ks->fs->device));
    return commit_fk20_compressed_dev(fk, d_polys, n, batch, (uint8_t*)d_commitments48, (uint8_t*)d_proofs48, (cudaStream_t)cuda_stream);
}
extern "C" int b200_commit_fk20_batch_compressed(b200_fk* fk, const uint64_t* polys, size_t n, size_t batch, uint8_t* commitments48,
                                                 uint8_t* proofs48) {
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fk->ks->fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf dp, dc, dout;
    CKS(dp.alloc(batch * n * 32, st)); CKS(dc.alloc(batch * 48, st)); CKS(dout.alloc(batch * n * 48, st));
    CK(cudaMemcpyAsync(dp.p, polys, batch * n * 32, cudaMemcpyHostToDevice, st));
    CKS(commit_fk20_compressed_dev(fk, dp.p, n, batch, dc.as<uint8_t>(), dout.as<uint8_t>(), st));
    CK(cudaMemcpyAsync(commitments48, dc.p, batch * 48, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(proofs48, dout.p, batch * n * 48, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
// ToCompressedG1 over an array of ABI points resident in HBM
extern "C" int b200_g1_compress_dev(const void* d_points, size_t n, void* d_out48, void* cuda_stream) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    if (n == 0) return B200_OK;
    CK(cudaSetDevice(g_device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DevBuf m;
    CKS(m.alloc(n * sizeof(G1J), st));
    launch_g1_from_abi((const uint64_t*)d_points, m.as<G1J>(), n, st);
    launch_g1_compress(m.as<G1J>(), (uint8_t*)d_out48, n, 1, 1, n, 0, 0, st);
    return check_launches();
}
extern "C" int b200_g1_to_compressed_batch(const uint64_t* points, size_t n, uint8_t* out48) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    if (n == 0) return B200_OK;
    CK(cudaSetDevice(g_device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    DevBuf in, out;
    CKS(in.alloc(n * 144, st)); CKS(out.alloc(n * 48, st));
    CK(cudaMemcpyAsync(in.p, points, n * 144, cudaMemcpyHostToDevice, st));
    CKS(b200_g1_compress_dev(in.p, n, out.p, st));
    CK(cudaMemcpyAsync(out48, out.p, n * 48, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return B200_OK;
}
// eth.BlobToKZGCommitment over a batch (eth/helpers.go:98-103 PolynomialToKZGCommitment after
// :264-273 BlobToPolynomial): `ks` holds the bit-reversed Lagrange setup (eth/globals.go:48), a blob is
// n x 32 little-endian bytes.  ok[b] = 0 and an all-zero commitment when blob b holds an element >= r.
extern "C" int b200_blob_to_kzg_commitment_batch(b200_ks* ks, const uint8_t* blobs, size_t n, size_t batch, uint8_t* out48, uint8_t* ok) {
    if (n > ks->n_g1) return B200_ERR_LEN_MISMATCH;
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(ks->fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    if (n == 0) { for (size_t b = 0; b < batch; b++) { memset(out48 + 48 * b, 0, 48); out48[48 * b] = 0xC0; ok[b] = 1; } return B200_OK; }
    DevBuf k, work, res, flags;
    CKS(k.alloc(batch * n * 32, st)); CKS(work.alloc(batch * n * sizeof(G1J), st)); CKS(res.alloc(batch * 48, st));
    CKS(flags.alloc(batch * 4, st));
    std::vector<uint32_t> h_ok(batch, 1u);
    CK(cudaMemcpyAsync(flags.p, h_ok.data(), batch * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(k.p, blobs, batch * n * 32, cudaMemcpyHostToDevice, st));
    launch_fr_check_canonical(k.as<uint64_t>(), n, batch, flags.as<uint32_t>(), st);
    const G1A* fb = nullptr;
    int fbw = 8;
    if (batch * n >= 4096) CKS(ks_fixed_base(ks, n, &fb, &fbw, st));
    CKS(dev_lincomb(ks->d_secret_g1, 0, k.as<Fr>(), 0, work.as<G1J>(), n, batch, st, fb, fbw));
    launch_g1_compress(work.as<G1J>(), res.as<uint8_t>(), 1, batch, 1, n, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(out48, res.p, batch * 48, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_ok.data(), flags.p, batch * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (size_t b = 0; b < batch; b++) {
        ok[b] = h_ok[b] ? 1 : 0;
        if (!h_ok[b]) memset(out48 + 48 * b, 0, 48);
    }
    return B200_OK;
}

// eth.ComputeKZGProof for a batch (eth/helpers.go:179-203): polys are evaluations on the bit-reversed domain
// (blob order), z one challenge per polynomial; proofs48 = ToCompressedG1(LinCombG1(lagrange, quotient)),
// y = f(z) (canonical, may be NULL).  ok[b] = 0 for "invalid z challenge" (z in the domain) or an input >= r.
extern "C" int b200_compute_kzg_proof_batch(b200_ks* ks, const uint64_t* polys, const uint64_t* z, size_t n, size_t batch,
                                            uint8_t* proofs48, uint64_t* y_out, uint8_t* ok) {
    b200_fs* fs = ks->fs;
    if (n > fs->max_width) return B200_ERR_TOO_LARGE;
    if (!is_pow2(n) || n < 16) return B200_ERR_NOT_POW2;
    if (n > ks->n_g1) return B200_ERR_LEN_MISMATCH;          // "polynomial has invalid length"
    if (batch == 0) return B200_OK;
    CK(cudaSetDevice(fs->device));
    StreamLease lease; CKS(lease.acquire()); cudaStream_t st = lease.st;
    const unsigned logn = log2u(n);
    DevBuf f, dz, inv, part, ym, yc, q, work, res, flags;
    CKS(f.alloc(batch * n * 32, st)); CKS(dz.alloc(batch * 32, st)); CKS(inv.alloc(batch * n * 32, st));
    CKS(part.alloc(batch * (n / 16) * 32, st)); CKS(ym.alloc(batch * 32, st)); CKS(yc.alloc(batch * 32, st));
    CKS(q.alloc(batch * n * 32, st)); CKS(work.alloc(batch * n * sizeof(G1J), st)); CKS(res.alloc(batch * 48, st));
    CKS(flags.alloc(batch * 4, st));
    std::vector<uint32_t> h_ok(batch, 1u);
    CK(cudaMemcpyAsync(flags.p, h_ok.data(), batch * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(f.p, polys, batch * n * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dz.p, z, batch * 32, cudaMemcpyHostToDevice, st));
    launch_fr_check_canonical(f.as<uint64_t>(), n, batch, flags.as<uint32_t>(), st);
    launch_fr_check_canonical(dz.as<uint64_t>(), 1, batch, flags.as<uint32_t>(), st);
    launch_eval_form_quotient(fs->dom, f.as<uint64_t>(), dz.as<uint64_t>(), logn, batch, 1, fr_inv_of_u64(n), inv.as<Fr>(), part.as<Fr>(),
                              ym.as<Fr>(), yc.as<uint64_t>(), q.as<uint64_t>(), flags.as<uint32_t>(), st);
    const G1A* fb = nullptr;
    int fbw = 8;
    if (batch * n >= 4096) CKS(ks_fixed_base(ks, n, &fb, &fbw, st));
    CKS(dev_lincomb(ks->d_secret_g1, 0, q.as<Fr>(), 0, work.as<G1J>(), n, batch, st, fb, fbw));
    launch_g1_compress(work.as<G1J>(), res.as<uint8_t>(), 1, batch, 1, n, 0, 0, st);
    CKS(check_launches());
    CK(cudaMemcpyAsync(proofs48, res.p, batch * 48, cudaMemcpyDeviceToHost, st));
    if (y_out) CK(cudaMemcpyAsync(y_out, yc.p, batch * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_ok.data(), flags.p, batch * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (size_t b = 0; b < batch; b++) {
        ok[b] = h_ok[b] ? 1 : 0;
        if (!h_ok[b]) { memset(proofs48 + 48 * b, 0, 48); if (y_out) memset(y_out + 4 * b, 0, 32); }
    }
    return B200_OK;
}

// ------------------------------------------------------------------------------ profiling
namespace b200 {
std::atomic<bool> g_prof_on{false};
struct ProfRec { int cat; cudaEvent_t e0, e1; bool closed; };
static std::vector<ProfRec> g_prof_recs;
static std::mutex g_prof_mu;
long prof_begin_event(int cat, cudaStream_t st) {
    ProfRec r; r.cat = cat; r.closed = false;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { cudaGetLastError(); return -1; }
    cudaEventRecord(r.e0, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(r);
    return (long)g_prof_recs.size() - 1;
}
void prof_end_event(long rec, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (rec < 0 || (size_t)rec >= g_prof_recs.size()) return;     // a profile_begin/end in between dropped the record
    cudaEventRecord(g_prof_recs[rec].e1, st);
    g_prof_recs[rec].closed = true;
}
}  // namespace b200
extern "C" int b200_profile_begin(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof_recs.clear();
    g_prof_on = true;
    return B200_OK;
}
extern "C" int b200_profile_end(double* ms_per_class, uint64_t* launches_per_class) {
    g_prof_on = false;
    CK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int c = 0; c < PROF_NCAT; c++) { ms_per_class[c] = 0; launches_per_class[c] = 0; }
    for (auto& r : g_prof_recs) {
        float ms = 0;
        if (r.closed && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { ms_per_class[r.cat] += ms; launches_per_class[r.cat]++; }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    cudaGetLastError();
    g_prof_recs.clear();
    return B200_OK;
}

// ------------------------------------------------------------------------------ self test / probes
extern "C" int b200_selftest_field(size_t n, uint64_t seed, uint64_t* mismatches) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    CK(cudaSetDevice(g_device));
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, 17 * 8));
    CK(cudaMemset(d, 0, 17 * 8));
    G1J* scratch = nullptr;
    CK(cudaMalloc(&scratch, (n ? 4 * n : 2) * sizeof(G1J)));
    launch_selftest(n, seed, d, scratch, nullptr);
    int rc = check_launches();
    cudaDeviceSynchronize();
    cudaFree(scratch);
    {   // host-recoded programs (both modes) for pseudo-random scalars and a few structured ones
        const size_t np = 256;
        std::vector<ScalarProgram> progs(np);
        std::vector<Fr> ks(np);
        uint64_t s = seed ^ 0xD1B54A32D192ED03ULL;
        for (size_t i = 0; i < np; i++) {
            Fr k;
            for (int j = 0; j < 8; j++) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; k.l[j] = (uint32_t)(s >> 32); }
            k.l[7] &= 0x3fffffffu;
            if (i < 32) k = fe_from_mont(fe_to_mont(fr_scale2_root_canon((unsigned)i)));   // the 2-adic roots of unity
            if (i == 32) { k = Fr::zero(); k.l[0] = 1; }
            if (i == 33) k = Fr::zero();
            ks[i] = k;
            make_scalar_program(&progs[i], k, (int)(i & 1));
        }
        ScalarProgram* dp = nullptr; Fr* dk = nullptr;
        CK(cudaMalloc(&dp, np * sizeof(ScalarProgram))); CK(cudaMalloc(&dk, np * sizeof(Fr)));
        CK(cudaMemcpy(dp, progs.data(), np * sizeof(ScalarProgram), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dk, ks.data(), np * sizeof(Fr), cudaMemcpyHostToDevice));
        launch_selftest_programs(np, dp, dk, d, nullptr);
        if (!rc) rc = check_launches();
        cudaDeviceSynchronize();
        cudaFree(dp); cudaFree(dk);
    }
    unsigned long long h[17] = {0};
    if (!rc && cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) { g_cuda_err = cudaGetErrorString(cudaGetLastError()); rc = B200_ERR_CUDA; }
    cudaFree(d);
    for (int i = 0; i < 17; i++) mismatches[i] = h[i];
    return rc;
}
// Fp multiplication throughput probe: `threads` lanes x `iters` x 2 dependent Montgomery products;
// returns the elapsed milliseconds (integer-pipe roofline denominator for the G1 kernels).
extern "C" int b200_probe_fp_mul(size_t threads, int iters, float* ms) {
    if (b200_device_count() == 0) return B200_ERR_NO_DEVICE;
    CK(cudaSetDevice(g_device));
    uint32_t* d = nullptr;
    CK(cudaMalloc(&d, 128));
    uint32_t h[32];
    for (int i = 0; i < 32; i++) h[i] = 0x9E3779B9u * (i + 1);
    CK(cudaMemcpy(d, h, 128, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch_fp_mul_probe(d, threads, 8, nullptr);   // warm-up
    CK(cudaEventRecord(e0, nullptr));
    launch_fp_mul_probe(d, threads, iters, nullptr);
    CK(cudaEventRecord(e1, nullptr));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return check_launches();
}
