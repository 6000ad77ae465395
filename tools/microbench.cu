// Pipe-cost microbenchmarks for sm_100a (B200): how many issue cycles per warp does each
// integer-multiply instruction form cost on one SM sub-partition?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
// Every kernel keeps ILP independent accumulators per thread; the multiplicand of accumulator k
// is taken from accumulator k+1 (so nothing is loop-invariant or warp-uniform).  Check the SASS
// (cuobjdump -sass) before trusting a line: ptxas rewrites naive loops.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__constant__ uint32_t c_b[64];

#define ILP 8
#define UNROLL 8
#define NB ((k + 1) & (ILP - 1))
#define KERNEL(NAME, DECL, INIT, BODY, FINI)                                                   \
    __global__ void __launch_bounds__(256) NAME(const uint32_t* __restrict__ in, uint32_t* out, int iters) { \
        const uint32_t* my = in + (threadIdx.x & 63) * 32;                                     \
        uint32_t a = my[30], b = my[31];                                                       \
        DECL;                                                                                  \
        _Pragma("unroll") for (int k = 0; k < ILP; k++) { INIT; }                              \
        _Pragma("unroll 1") for (int it = 0; it < iters; it++) {                               \
            _Pragma("unroll") for (int u = 0; u < UNROLL; u++) {                               \
                _Pragma("unroll") for (int k = 0; k < ILP; k++) { BODY; }                      \
            }                                                                                  \
        }                                                                                      \
        uint32_t r = a ^ b;                                                                    \
        _Pragma("unroll") for (int k = 0; k < ILP; k++) { FINI; }                              \
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;                                \
    }
#define X64(v) ((uint32_t)(v) ^ (uint32_t)((v) >> 32))
#define INIT64(v, o) v[k] = ((uint64_t)my[o + k] << 32) | my[o + 8 + k]

// IMAD.WIDE.U32 Rd64 = x*b + Rc64 (no carry predicates), all-register operands
KERNEL(k_wide, uint64_t acc[ILP], INIT64(acc, 0),
       asm volatile("{\n\t.reg .u32 xl, xh;\n\tmov.b64 {xl, xh}, %1;\n\tmad.wide.u32 %0, xl, %2, %0;\n\t}"
                    : "+l"(acc[k]) : "l"(acc[NB]), "r"(b)),
       r ^= X64(acc[k]))
// pure product, no addend: IMAD.WIDE.U32 Rd64 = x*b + RZ
KERNEL(k_mulwide, uint64_t acc[ILP], INIT64(acc, 0),
       asm volatile("{\n\t.reg .u32 xl, xh;\n\tmov.b64 {xl, xh}, %1;\n\tor.b32 xl, xl, 1;\n\tmul.wide.u32 %0, xl, %2;\n\t}"
                    : "=l"(acc[k]) : "l"(acc[NB]), "r"(b)),
       r ^= X64(acc[k]))
// same with the multiplier in constant memory (c[bank][off] / uniform-register operand)
KERNEL(k_wide_const, uint64_t acc[ILP], INIT64(acc, 0),
       asm volatile("{\n\t.reg .u32 xl, xh;\n\tmov.b64 {xl, xh}, %1;\n\tmad.wide.u32 %0, xl, %2, %0;\n\t}"
                    : "+l"(acc[k]) : "l"(acc[NB]), "r"(c_b[k])),
       r ^= X64(acc[k]))
// carry-OUT only: IMAD.WIDE.U32 Rd, Pout = ... ; IADD3.X captures the carry
KERNEL(k_wide_cout, uint32_t lo[ILP]; uint32_t hi[ILP]; uint32_t t[ILP], (lo[k] = my[k], hi[k] = my[8 + k], t[k] = 0),
       asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                    : "+r"(lo[k]), "+r"(hi[k]), "+r"(t[k]) : "r"(hi[NB]), "r"(b)),
       r ^= lo[k] ^ hi[k] ^ t[k])
// carry-IN only: IADD3 produces a carry consumed by IMAD.WIDE.U32.X (no carry-out)
KERNEL(k_wide_cin, uint32_t lo[ILP]; uint32_t hi[ILP]; uint32_t t[ILP], (lo[k] = my[k], hi[k] = my[8 + k], t[k] = my[16 + k]),
       asm volatile("add.cc.u32 %2, %2, %3;\n\tmadc.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.u32 %1, %3, %4, %1;"
                    : "+r"(lo[k]), "+r"(hi[k]), "+r"(t[k]) : "r"(hi[NB]), "r"(b)),
       r ^= lo[k] ^ hi[k] ^ t[k])
// 4-link carry chain (the production idiom): 4 IMAD.WIDE.U32(.X) + IADD3.X
KERNEL(k_chain4, uint32_t e[ILP][9], for (int q = 0; q < 9; q++) e[k][q] = my[(k + q) & 31],
       asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                    "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                    "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                    "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                    : "+r"(e[k][0]), "+r"(e[k][1]), "+r"(e[k][2]), "+r"(e[k][3]), "+r"(e[k][4]), "+r"(e[k][5]),
                      "+r"(e[k][6]), "+r"(e[k][7]), "+r"(e[k][8])
                    : "r"(a), "r"(b), "r"(a ^ 0x5555u), "r"(b ^ 0x3333u), "r"(e[NB][1])),
       for (int q = 0; q < 9; q++) r ^= e[k][q])
// IMAD (32-bit low) and IMAD.HI.U32
KERNEL(k_imad_lo, uint32_t acc[ILP], acc[k] = my[k],
       asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[NB]), "r"(b)), r ^= acc[k])
KERNEL(k_imad_hi, uint32_t acc[ILP], acc[k] = my[k],
       asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[NB]), "r"(b)), r ^= acc[k])
// ALU pipe: 64-bit add = IADD3 + IADD3.X
KERNEL(k_iadd64, uint32_t lo[ILP]; uint32_t hi[ILP], (lo[k] = my[k], hi[k] = my[8 + k]),
       asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[k]), "+r"(hi[k]) : "r"(lo[NB]), "r"(hi[NB])),
       r ^= lo[k] ^ hi[k])
KERNEL(k_lop3, uint32_t acc[ILP], acc[k] = my[k],
       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[k]) : "r"(acc[NB]), "r"(b)), r ^= acc[k])
KERNEL(k_shf, uint32_t acc[ILP], acc[k] = my[k],
       asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(acc[k]) : "r"(acc[NB])), r ^= acc[k])
// FP pipes
KERNEL(k_dfma, double acc[ILP]; double db = (double)b * 1e-10, acc[k] = (double)my[k] * 1e-9,
       asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[k]) : "d"(acc[NB]), "d"(db)),
       r ^= (uint32_t)__double_as_longlong(acc[k]))
KERNEL(k_ffma, float acc[ILP]; float fb = (float)b * 1e-10f, acc[k] = (float)my[k] * 1e-9f,
       asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k]) : "f"(acc[NB]), "f"(fb)), r ^= __float_as_uint(acc[k]))
// co-issue: 1 IMAD.WIDE + one 64-bit add (2 ALU) per body
KERNEL(k_mix_wide_iadd, uint64_t acc[ILP]; uint32_t lo[ILP]; uint32_t hi[ILP], (INIT64(acc, 0), lo[k] = my[16 + k], hi[k] = my[20 + k]),
       asm volatile("{\n\t.reg .u32 xl, xh;\n\tmov.b64 {xl, xh}, %3;\n\tmad.wide.u32 %0, xl, %4, %0;\n\t"
                    "add.cc.u32 %1, %1, %5;\n\taddc.u32 %2, %2, %6;\n\t}"
                    : "+l"(acc[k]), "+r"(lo[k]), "+r"(hi[k]) : "l"(acc[NB]), "r"(b), "r"(lo[NB]), "r"(hi[NB])),
       r ^= X64(acc[k]) ^ lo[k] ^ hi[k])
// co-issue: 1 IMAD.WIDE + 1 DFMA per body
KERNEL(k_mix_wide_dfma, uint64_t acc[ILP]; double dacc[ILP]; double db = (double)b * 1e-10, (INIT64(acc, 0), dacc[k] = (double)my[16 + k] * 1e-9),
       asm volatile("{\n\t.reg .u32 xl, xh;\n\tmov.b64 {xl, xh}, %2;\n\tmad.wide.u32 %0, xl, %3, %0;\n\t"
                    "fma.rn.f64 %1, %4, %5, %1;\n\t}"
                    : "+l"(acc[k]), "+d"(dacc[k]) : "l"(acc[NB]), "r"(b), "d"(dacc[NB]), "d"(db)),
       r ^= X64(acc[k]) ^ (uint32_t)__double_as_longlong(dacc[k]))
// co-issue: 1 IMAD.WIDE + 1 FFMA per body (fmaheavy + fmalite?)
KERNEL(k_mix_wide_ffma, uint64_t acc[ILP]; float facc[ILP]; float fb = (float)b * 1e-10f, (INIT64(acc, 0), facc[k] = (float)my[16 + k] * 1e-9f),
       asm volatile("{\n\t.reg .u32 xl, xh;\n\tmov.b64 {xl, xh}, %2;\n\tmad.wide.u32 %0, xl, %3, %0;\n\t"
                    "fma.rn.f32 %1, %4, %5, %1;\n\t}"
                    : "+l"(acc[k]), "+f"(facc[k]) : "l"(acc[NB]), "r"(b), "f"(facc[NB]), "f"(fb)),
       r ^= X64(acc[k]) ^ __float_as_uint(facc[k]))
// co-issue: 1 chained IMAD.WIDE.X + 1 DFMA
KERNEL(k_mix_chain_dfma, uint32_t e[ILP][9]; double dacc[ILP]; double db = (double)b * 1e-10,
       (dacc[k] = (double)my[16 + k] * 1e-9); for (int q = 0; q < 9; q++) e[k][q] = my[(k + q) & 31],
       asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                    "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                    "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                    "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;\n\t"
                    "fma.rn.f64 %14, %15, %16, %14;\n\tfma.rn.f64 %14, %15, %16, %14;\n\t"
                    "fma.rn.f64 %14, %15, %16, %14;\n\tfma.rn.f64 %14, %15, %16, %14;"
                    : "+r"(e[k][0]), "+r"(e[k][1]), "+r"(e[k][2]), "+r"(e[k][3]), "+r"(e[k][4]), "+r"(e[k][5]),
                      "+r"(e[k][6]), "+r"(e[k][7]), "+r"(e[k][8])
                    : "r"(a), "r"(b), "r"(a ^ 0x5555u), "r"(b ^ 0x3333u), "r"(e[NB][1]), "d"(dacc[k]), "d"(dacc[NB]), "d"(db)),
       for (int q = 0; q < 9; q++) r ^= e[k][q]; r ^= (uint32_t)__double_as_longlong(dacc[k]))

typedef void (*kern_t)(const uint32_t*, uint32_t*, int);
struct Entry { const char* name; kern_t fn; int instr_per_body; const char* note; };

int main() {
    const int threads = 256, blocks = 148 * 8, iters = 1024;
    uint32_t *d_in, *d_out;
    std::vector<uint32_t> h(64 * 32);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)(0x9e3779b9u * (uint32_t)(i + 1)) ^ (uint32_t)(i * 2654435761u >> 7) | 1u;
    cudaMalloc(&d_in, h.size() * 4); cudaMalloc(&d_out, (size_t)threads * blocks * 4);
    cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_b, h.data(), 256);
    Entry tab[] = {
        {"imad_wide (64b addend, reg)", k_wide, 1, "IMAD.WIDE.U32"},
        {"mul.wide (no addend) [+1 LOP3]", k_mulwide, 1, "IMAD.WIDE.U32 .., RZ"},
        {"imad_wide (64b addend, const mult)", k_wide_const, 1, "IMAD.WIDE.U32 UR/c[]"},
        {"imad_wide carry-out only", k_wide_cout, 1, "+1 IADD3.X per IMAD"},
        {"imad_wide.X carry-in only", k_wide_cin, 1, "+1 IADD3 per IMAD"},
        {"chain4 (cin+cout)", k_chain4, 4, "per IMAD, +1/4 IADD3.X"},
        {"imad lo32", k_imad_lo, 1, "IMAD"},
        {"imad.hi.u32", k_imad_hi, 1, "IMAD.HI.U32"},
        {"iadd3+iadd3.x (64-bit add)", k_iadd64, 2, "per ALU instr"},
        {"lop3", k_lop3, 1, "LOP3"},
        {"shf", k_shf, 1, "SHF"},
        {"dfma", k_dfma, 1, "DFMA"},
        {"ffma", k_ffma, 1, "FFMA"},
        {"mix: imad_wide + 64-bit add", k_mix_wide_iadd, 1, "per body (1 IMAD + 2 ALU)"},
        {"mix: imad_wide + dfma", k_mix_wide_dfma, 1, "per body (1 IMAD + 1 DFMA)"},
        {"mix: imad_wide + ffma", k_mix_wide_ffma, 1, "per body (1 IMAD + 1 FFMA)"},
        {"mix: chain4 + 4 dfma", k_mix_chain_dfma, 4, "per IMAD (1 DFMA each)"},
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("grid %d x %d, ILP %d\n", blocks, threads, ILP);
    printf("%-40s %10s %12s  %s\n", "variant", "Tinstr/s", "cyc/warp-inst", "note (cycles per SM sub-partition at 1965 MHz)");
    for (auto& e : tab) {
        e.fn<<<blocks, threads>>>(d_in, d_out, 32);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); e.fn<<<blocks, threads>>>(d_in, d_out, iters); cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double instr = (double)threads * blocks * iters * ILP * UNROLL * e.instr_per_body;
        double per_s = instr / (best * 1e-3);
        double cyc = 1.965e9 / (per_s / 32.0 / (148.0 * 4.0));
        printf("%-40s %10.3f %12.3f  %s\n", e.name, per_s / 1e12, cyc, e.note);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
    return 0;
}
