#include <cstdint>
#include <cstdio>
struct Acc { double c[32]; };
__device__ __forceinline__ void tm_park(unsigned taddr, const Acc &a) {
    unsigned r[64];
#pragma unroll
    for (int i = 0; i < 32; ++i) { r[2 * i] = (unsigned)__double2loint(a.c[i]); r[2 * i + 1] = (unsigned)__double2hiint(a.c[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,"
                 "%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63,%64};"
                 :: "r"(taddr),
                 "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),
                 "r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]),
                 "r"(r[32]),"r"(r[33]),"r"(r[34]),"r"(r[35]),"r"(r[36]),"r"(r[37]),"r"(r[38]),"r"(r[39]),"r"(r[40]),"r"(r[41]),"r"(r[42]),"r"(r[43]),"r"(r[44]),"r"(r[45]),"r"(r[46]),"r"(r[47]),
                 "r"(r[48]),"r"(r[49]),"r"(r[50]),"r"(r[51]),"r"(r[52]),"r"(r[53]),"r"(r[54]),"r"(r[55]),"r"(r[56]),"r"(r[57]),"r"(r[58]),"r"(r[59]),"r"(r[60]),"r"(r[61]),"r"(r[62]),"r"(r[63])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_unpark(unsigned taddr, Acc &a) {
    unsigned r[64];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                 "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]),
                 "=r"(r[32]),"=r"(r[33]),"=r"(r[34]),"=r"(r[35]),"=r"(r[36]),"=r"(r[37]),"=r"(r[38]),"=r"(r[39]),"=r"(r[40]),"=r"(r[41]),"=r"(r[42]),"=r"(r[43]),"=r"(r[44]),"=r"(r[45]),"=r"(r[46]),"=r"(r[47]),
                 "=r"(r[48]),"=r"(r[49]),"=r"(r[50]),"=r"(r[51]),"=r"(r[52]),"=r"(r[53]),"=r"(r[54]),"=r"(r[55]),"=r"(r[56]),"=r"(r[57]),"=r"(r[58]),"=r"(r[59]),"=r"(r[60]),"=r"(r[61]),"=r"(r[62]),"=r"(r[63])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) a.c[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__global__ void __launch_bounds__(128, 4) k(double *out, const double *in) {
    __shared__ unsigned s_tm;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(&s_tm)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = s_tm + ((unsigned)(32 * w) << 16);
    Acc a, b;
    for (int i = 0; i < 32; ++i) { a.c[i] = in[(blockIdx.x * 128 + threadIdx.x) * 64 + i]; b.c[i] = in[(blockIdx.x * 128 + threadIdx.x) * 64 + 32 + i]; }
    tm_park(base, a);
    tm_park(base + 64, b);
    Acc c, d;
    tm_unpark(base + 64, c);
    tm_unpark(base, d);
    for (int i = 0; i < 32; ++i) { out[(blockIdx.x * 128 + threadIdx.x) * 64 + i] = d.c[i]; out[(blockIdx.x * 128 + threadIdx.x) * 64 + 32 + i] = c.c[i]; }
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tm), "n"(128) : "memory");
}
int main() {
    const int nb = 148 * 8, n = nb * 128 * 64;
    double *in, *out;
    cudaMalloc(&in, n * 8); cudaMalloc(&out, n * 8);
    double *h = new double[n];
    for (int i = 0; i < n; ++i) h[i] = i * 0.5 + 1.0 / (i + 1);
    cudaMemcpy(in, h, n * 8, cudaMemcpyHostToDevice);
    cudaMemset(out, 0, n * 8);
    k<<<nb, 128>>>(out, in);
    cudaError_t e = cudaDeviceSynchronize();
    double *g = new double[n];
    cudaMemcpy(g, out, n * 8, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < n; ++i) bad += g[i] != h[i];
    printf("tmem park/unpark: %s, mismatches %ld of %d\n", cudaGetErrorString(e), bad, n);
    return bad != 0 || e != cudaSuccess;
}
