// Developer probe: which cp.async.bulk.tensor configurations does this box accept?  Each case runs in a
// child process so that a faulting case does not poison the others.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_probe tools/probes/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe4d(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int use_global,
                        int c0, int c1, int c2, int c3, uint32_t bytes, float *out, int nout)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t mbar = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap *tm = use_global ? gmap : &tmap;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(smem)),
            "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(mbar)
            : "memory");
    }
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(mbar), "r"(0u) : "memory");
    __syncthreads();
    const float *s = reinterpret_cast<const float *>(smem);
    for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = ok ? s[i] : -12345.0f;
}

static int run_case(const char *name, int W, int H, int C, int B, int bw, int bh, int bc, int c0, int c1,
                    int use_global)
{
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) {
        printf("%s: no entry point\n", name);
        return 2;
    }
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const size_t n = (size_t)W * H * C * B;
    float *hx = (float *)malloc(n * 4);
    for (size_t i = 0; i < n; ++i) hx[i] = (float)(i % 100003);
    float *dx = nullptr, *dout = nullptr;
    cudaMalloc(&dx, n * 4);
    cudaMemcpy(dx, hx, n * 4, cudaMemcpyHostToDevice);
    const int nout = bw * bh * bc;
    cudaMalloc(&dout, (size_t)nout * 4);
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("%s: encode rc=%d\n", name, (int)rc); return 3; }
    CUtensorMap *gmap = nullptr;
    cudaMalloc(&gmap, sizeof(CUtensorMap));
    cudaMemcpy(gmap, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    const size_t smem = (size_t)nout * 4 + 1024;
    cudaFuncSetAttribute(probe4d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe4d<<<1, 128, smem>>>(tm, gmap, use_global, c0, c1, 0, 0, (uint32_t)(nout * 4), dout, nout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: KERNEL ERROR %s\n", name, cudaGetErrorString(e)); return 4; }
    float *ho = (float *)malloc((size_t)nout * 4);
    cudaMemcpy(ho, dout, (size_t)nout * 4, cudaMemcpyDeviceToHost);
    // check a few elements: out[(c*bh + y)*bw + x] == x[(c*H + c1+y)*W + c0+x] or 0 outside
    int bad = 0;
    for (int c = 0; c < bc; ++c)
        for (int y = 0; y < bh; ++y)
            for (int x = 0; x < bw; ++x) {
                const int gx = c0 + x, gy = c1 + y;
                const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? hx[((size_t)c * H + gy) * W + gx] : 0.0f;
                if (ho[((size_t)c * bh + y) * bw + x] != want) ++bad;
            }
    printf("%s: ok, %d of %d elements wrong (first %.1f)\n", name, bad, nout, ho[0]);
    return bad ? 5 : 0;
}

int main(int argc, char **argv)
{
    struct Case { const char *name; int W, H, C, B, bw, bh, bc, c0, c1, use_global; };
    const Case cases[] = {
        {"A box 32x8x4 at (0,0), param map", 128, 64, 32, 2, 32, 8, 4, 0, 0, 0},
        {"C box 36x10x4 at (-1,-1), param map", 128, 64, 32, 2, 36, 10, 4, -1, -1, 0},
        {"G box 32x10x32 at (0,-1), param map", 128, 64, 32, 2, 32, 10, 32, 0, -1, 0},
        {"I box 32x10x32 at (-1,-1)", 128, 64, 32, 2, 32, 10, 32, -1, -1, 0},
        {"J box 4x10x32 at (31,-1)", 128, 64, 32, 2, 4, 10, 32, 31, -1, 0},
        {"K box 8x10x32 at (31,-1)", 128, 64, 32, 2, 8, 10, 32, 31, -1, 0},
        {"L box 40x10x32 at (-4,-1)", 128, 64, 32, 2, 40, 10, 32, -4, -1, 0},
        {"M box 48x10x32 at (-1,-1)", 128, 64, 32, 2, 48, 10, 32, -1, -1, 0},
        {"N box 64x10x32 at (-1,-1)", 128, 64, 32, 2, 64, 10, 32, -1, -1, 0},
        {"O box 36x10x8 at (0,0)", 128, 64, 32, 2, 36, 10, 8, 0, 0, 0},
        {"P box 36x2x2 at (0,0)", 128, 64, 32, 2, 36, 2, 2, 0, 0, 0},
        {"Q box 64x10x32 at (-1,-1), W=1920 H=1080", 1920, 1080, 32, 1, 64, 10, 32, -1, -1, 0},
    };
    const int ncases = (int)(sizeof(cases) / sizeof(cases[0]));
    if (argc > 1) {
        const Case &c = cases[atoi(argv[1])];
        return run_case(c.name, c.W, c.H, c.C, c.B, c.bw, c.bh, c.bc, c.c0, c.c1, c.use_global);
    }
    for (int i = 0; i < ncases; ++i) {
        fflush(stdout);
        pid_t pid = fork();
        if (pid == 0) {
            char idx[8];
            snprintf(idx, sizeof idx, "%d", i);
            execl(argv[0], argv[0], idx, (char *)nullptr);
            _exit(99);
        }
        int st = 0;
        waitpid(pid, &st, 0);
    }
    return 0;
}
