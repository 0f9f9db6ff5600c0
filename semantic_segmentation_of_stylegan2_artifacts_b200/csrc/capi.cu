// C-ABI glue: error state, launch counter, GEMM backend dispatch.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace msu {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static thread_local int g_last_backend = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static std::atomic<int> g_deterministic{-1};
int deterministic_mode() {
    int v = g_deterministic.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("MSU_DETERMINISTIC");
        v = (e != nullptr && atoi(e) != 0) ? 1 : 0;
        g_deterministic.store(v, std::memory_order_relaxed);
    }
    return v;
}
void set_deterministic_mode(int on) { g_deterministic.store(on ? 1 : 0, std::memory_order_relaxed); }
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int gemm_simt(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t M, int64_t N, int64_t K,
              float* splitk_ws, int64_t splitk_ws_elems, cudaStream_t st);
// returns 1 if the pattern is not supported by the tcgen05 path (caller falls back to SIMT), 0 ok, else error
int gemm_tc(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t M, int64_t N, int64_t K,
            float* splitk_ws, int64_t splitk_ws_elems, cudaStream_t st);
int wgrad_tc(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t I, int64_t J, int64_t T,
             float* ws, int64_t ws_elems, cudaStream_t st, int* fused_colsum);

}  // namespace msu

using namespace msu;

extern "C" int msu_gemm(const MsuOperand* A, const MsuOperand* B, const MsuEpilogue* E, int64_t M, int64_t N, int64_t K,
                        float* splitk_ws, int64_t splitk_ws_elems, int backend, void* stream) {
    MSU_REQUIRE(A && B && E && A->ptr && B->ptr && (E->C || E->lnd_w), "msu_gemm: null pointer");
    MSU_REQUIRE(M >= 0 && N > 0 && K > 0, "msu_gemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    g_last_backend = 0;
    const bool wgrad = (A->orient == 1 && B->orient == 1);
    MSU_REQUIRE(E->colsum == nullptr || (wgrad && E->out_f32), "msu_gemm: colsum is defined for weight-gradient GEMMs (orient 1 operands, fp32 output) only");
    int rc = 1, fused = 0;
    if (backend == 0) {
        rc = wgrad ? wgrad_tc(A, B, E, M, N, K, splitk_ws, splitk_ws_elems, st, &fused)
                   : gemm_tc(A, B, E, M, N, K, splitk_ws, splitk_ws_elems, st);
        if (rc == 0) g_last_backend = 1;
        else if (rc != 1) return rc;
    }
    MSU_REQUIRE(!(rc == 1 && E->lnd_w != nullptr), "msu_gemm: the fused LayerNorm + dot epilogue needs the tcgen05 TMA-store path "
                "(bf16, unmapped output, one N tile of a multiple of 32 columns, no other fused operand)");
    if (rc == 1) {
        rc = gemm_simt(A, B, E, M, N, K, splitk_ws, splitk_ws_elems, st);
        if (rc != 0) return rc;
    }
    if (E->colsum != nullptr && !fused) {
        // bias gradient not fused by the kernel above: column sums of A read as [K rows (tokens), M columns]
        // (stream ordered after the split-K reduce, so the workspace can be shared)
        MsuOperand X = *A;
        X.orient = 0;
        return msu_colsum(&X, K, M, E->colsum, 0, splitk_ws, splitk_ws_elems, stream);
    }
    return 0;
}

extern "C" int msu_version(void) { return 100; }
extern "C" int msu_struct_size(int which) { return which == 0 ? (int)sizeof(MsuOperand) : (int)sizeof(MsuEpilogue); }
extern "C" const char* msu_last_error_string(void) { return g_err; }
extern "C" long long msu_launch_count(void) { return g_launches.load(); }
extern "C" int msu_set_deterministic(int on) {
    const int prev = msu::deterministic_mode();
    msu::set_deterministic_mode(on);
    return prev;
}
extern "C" int msu_last_gemm_backend(void) { return g_last_backend; }
