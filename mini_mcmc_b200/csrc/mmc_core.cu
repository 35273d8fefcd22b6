// Library plumbing: error strings, device selection, and the host-side `init` routines
// (src/core.rs:394-435) that must stay bit-compatible with rand 0.9 SmallRng + rand_distr 0.5
// StandardNormal so that init_det()/init_with_seed() keep producing the reference's starting points.
#include <stdarg.h>
#include <string.h>

#include <cmath>
#include <mutex>
#include <string>
#include <vector>

#include "mmc_common.cuh"

namespace mmc {

namespace {
struct NamedTarget { std::string name; CustomTargetEntry e; };
std::mutex g_custom_mutex;
std::vector<NamedTarget> &custom_targets() {
    static std::vector<NamedTarget> r;
    return r;
}
}  // namespace

int custom_target_register(const char *name, int dim, mmc_hmc_launch_fn hmc, mmc_nuts_launch_fn nuts, mmc_mh_launch_fn mh) {
    if (!name || dim <= 0 || (!hmc && !nuts && !mh)) {
        set_error("custom target registration: bad arguments");
        return MMC_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lock(g_custom_mutex);
    auto &r = custom_targets();
    size_t i = 0;
    for (; i < r.size(); ++i)
        if (r[i].name == name) break;
    if (i == r.size()) r.push_back({name, {}});
    if (r[i].e.dim != 0 && r[i].e.dim != dim) {
        set_error("custom target '%s' is already registered with dim %d (got %d)", name, r[i].e.dim, dim);
        return MMC_ERR_INVALID;
    }
    r[i].e.dim = dim;
    if (hmc) r[i].e.hmc = hmc;
    if (nuts) r[i].e.nuts = nuts;
    if (mh) r[i].e.mh = mh;
    return MMC_T_CUSTOM_BASE + (int)i;
}

bool custom_target_get(int kind, CustomTargetEntry *out, const char **name) {
    std::lock_guard<std::mutex> lock(g_custom_mutex);
    auto &r = custom_targets();
    const size_t idx = (size_t)(kind - MMC_T_CUSTOM_BASE);
    if (kind < MMC_T_CUSTOM_BASE || idx >= r.size()) return false;
    if (out) *out = r[idx].e;
    if (name) *name = r[idx].name.c_str();
    return true;
}

int custom_target_lookup(const char *name) {
    if (!name) return MMC_ERR_INVALID;
    std::lock_guard<std::mutex> lock(g_custom_mutex);
    auto &r = custom_targets();
    for (size_t i = 0; i < r.size(); ++i)
        if (r[i].name == name) return MMC_T_CUSTOM_BASE + (int)i;
    return MMC_ERR_INVALID;
}

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) in `%s` at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError)
        return MMC_ERR_NO_DEVICE;
    if (e == cudaErrorMemoryAllocation) return MMC_ERR_NOMEM;
    return MMC_ERR_CUDA;
}

static int g_sm_count = 0;

int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); libminimcmc has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return MMC_ERR_NO_DEVICE;
    }
    if (g_sm_count == 0) {
        int dev = 0;
        MMC_CUDA(cudaGetDevice(&dev));
        MMC_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    return MMC_OK;
}

int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

// ---------------------------------------------------------------- host RNG for init (rand 0.9 / rand_distr 0.5)
// Xoshiro256++ seeded through SplitMix64 and the 256-layer ziggurat; tables regenerated with the
// recurrence rand_distr's ziggurat_tables.py uses.
struct HostSmallRng {
    uint64_t s[4];
    explicit HostSmallRng(uint64_t state) {
        for (auto &w : s) {
            state += 0x9e3779b97f4a7c15ULL;
            uint64_t z = state;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            w = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
};

static double bits_to_f64(uint64_t frac52, int e) {
    const uint64_t b = frac52 | ((uint64_t)(1023 + e) << 52);
    double d;
    memcpy(&d, &b, 8);
    return d;
}

struct ZigNormal {
    double x[257], f[257];
    static constexpr double R = 3.654152885361008796, V = 0.00492867323399;
    ZigNormal() {
        auto pdf = [](double v) { return std::exp(-v * v / 2.0); };
        x[0] = V / pdf(R);
        x[1] = R;
        for (int i = 2; i < 256; ++i) x[i] = std::sqrt(-2.0 * std::log(V / x[i - 1] + pdf(x[i - 1])));
        x[256] = 0.0;
        for (int i = 0; i < 257; ++i) f[i] = pdf(x[i]);
    }
    double sample(HostSmallRng &rng) const {
        for (;;) {
            const uint64_t bits = rng.next();
            const int i = (int)(bits & 0xff);
            const double u = bits_to_f64(bits >> 12, 1) - 3.0;
            const double v = u * x[i];
            if (std::fabs(v) < x[i + 1]) return v;
            if (i == 0) {
                double tx = 1.0, ty = 0.0;
                while (-2.0 * ty < tx * tx) {
                    const double a = bits_to_f64(rng.next() >> 12, 0) - (1.0 - 0x1p-53);
                    const double b = bits_to_f64(rng.next() >> 12, 0) - (1.0 - 0x1p-53);
                    tx = std::log(a) / R;
                    ty = std::log(b);
                }
                return u < 0.0 ? tx - R : R - tx;
            }
            const double uf = (double)(rng.next() >> 11) * 0x1p-53;
            if (f[i + 1] + (f[i] - f[i + 1]) * uf < std::exp(-v * v / 2.0)) return v;
        }
    }
};

__global__ void init_positions_kernel(float *out, int64_t n, int64_t d, uint2 key, int64_t chain_offset) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (chain, group of 4 dims)
    const int64_t groups = (d + 3) / 4;
    if (idx >= n * groups) return;
    const int64_t c = idx / groups, j = idx % groups;
    const uint64_t gchain = (uint64_t)(c + chain_offset);
    const uint4 w = philox4x32_10(key, make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), kStepInit, (uint32_t)j));
    float nn[4];
    box_muller_f32(w.x, w.y, nn[0], nn[1]);
    box_muller_f32(w.z, w.w, nn[2], nn[3]);
    for (int k = 0; k < 4; ++k)
        if (4 * j + k < d) out[c * d + 4 * j + k] = nn[k];
}

}  // namespace mmc

extern "C" {

int mmc_version(void) { return MMC_VERSION; }

const char *mmc_last_error(void) { return mmc::g_err; }

int mmc_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        mmc::set_error("no CUDA device available (%s); libminimcmc has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return MMC_ERR_NO_DEVICE;
    }
    MMC_REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    MMC_CUDA(cudaSetDevice(device));
    mmc::g_sm_count = 0;
    return mmc::ensure_device();
}

int mmc_device_info(int *sm_count, char *name, int name_len) {
    int rc = mmc::ensure_device();
    if (rc) return rc;
    int dev = 0;
    MMC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MMC_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (name && name_len > 0) {
        strncpy(name, prop.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    return MMC_OK;
}

int mmc_init_positions(double *out_host, int64_t n, int64_t d, uint64_t seed) {
    MMC_REQUIRE(out_host && n >= 0 && d >= 0, "mmc_init_positions: bad arguments");
    static const mmc::ZigNormal zig;
    mmc::HostSmallRng rng(seed);
    for (int64_t i = 0; i < n * d; ++i) out_host[i] = zig.sample(rng);
    return MMC_OK;
}

int mmc_init_positions_dev(float *out_dev, int64_t n, int64_t d, uint64_t seed, int64_t chain_offset, void *stream) {
    int rc = mmc::ensure_device();
    if (rc) return rc;
    MMC_REQUIRE(out_dev && n > 0 && d > 0, "mmc_init_positions_dev: bad arguments");
    const int64_t total = n * ((d + 3) / 4);
    const int block = 256;
    mmc::init_positions_kernel<<<(unsigned)((total + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
        out_dev, n, d, mmc::seed_key(seed), chain_offset);
    MMC_CUDA(cudaGetLastError());
    return MMC_OK;
}

}  // extern "C"
