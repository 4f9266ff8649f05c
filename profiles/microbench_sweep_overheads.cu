// experiment: what do the non-streaming parts of the sweep cost? (not part of the product)
// base = microbench_sweep_access_pattern.cu with gathers; flags add, one at a time, the pieces the product row has on top:
//   NORM  : |dx|, psi = |x - z|, norm /= psi when psi > 1 (one fp64 division per row), warp + block reduction, one partial per block
//   PID   : 2-byte pattern id per row and the "hot pattern" branch (offsets from kernel parameters, else from a table)
//   VIEW  : a ~900-byte kernel parameter block (the product passes its whole array view by value)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
struct Big { const double *m, *b, *z; const uint16_t *pid; const int *table; double *part; size_t N; int hot[10]; uint32_t hotPid; char pad[760]; };
template <int NORM, int PID>
__global__ void __launch_bounds__(256) k_sweep(Big v, const double* __restrict__ xin, double* __restrict__ xout)
{
    __shared__ double sh[8];
    const size_t N = v.N;
    double norm = 0.;
    for (size_t i = blockIdx.x*256ull+threadIdx.x; i < N; i += (size_t)gridDim.x*256) {
        long off[10];
        if (PID) {
            const uint32_t p = v.pid[i];
            if (p == v.hotPid) { _Pragma("unroll") for (int c=0;c<10;++c) off[c] = v.hot[c]; }
            else { const int *t = v.table + p*10; _Pragma("unroll") for (int c=0;c<10;++c) off[c] = t[c]; }
        } else { _Pragma("unroll") for (int c=0;c<10;++c) off[c] = v.hot[c]; }
        double acc = __ldcs(v.b+i);
        #pragma unroll
        for (int c=0;c<10;++c) {
            double a = __ldcs(v.m + (size_t)c*N + i);
            long j = (long)i + off[c]; if (j<0) j=i; if (j>=(long)N) j=i;
            acc -= a*xin[j];
        }
        const double zz = __ldcs(v.z+i);
        const double xo = xin[i];
        if (NORM) {
            double nr = fabs(acc - xo);
            const double psi = fabs(acc - zz);
            if (psi > 1.) nr *= (1. / psi);
            norm += nr;
            xout[i] = acc;
        } else xout[i] = acc + zz*1e-30 + xo*1e-30;
    }
    if (NORM) {
        for (int o=16;o>0;o>>=1) norm += __shfl_down_sync(0xffffffffu, norm, o);
        if ((threadIdx.x&31)==0) sh[threadIdx.x>>5] = norm;
        __syncthreads();
        if (threadIdx.x < 32) { double t = threadIdx.x < 8 ? sh[threadIdx.x] : 0.; for (int o=4;o>0;o>>=1) t += __shfl_down_sync(0xffffffffu, t, o); if (threadIdx.x==0) v.part[blockIdx.x] = t; }
    }
}
template <int NORM, int PID> float run(Big v, double *x0, double *x1, int blocks)
{
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w=0;w<3;++w) k_sweep<NORM,PID><<<blocks,256>>>(v,x0,x1);
    cudaEventRecord(e0);
    for (int it=0; it<20; ++it) k_sweep<NORM,PID><<<blocks,256>>>(v,(it&1)?x1:x0,(it&1)?x0:x1);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1); return ms/20;
}
int main(){
    size_t R=1024,C=1024,L=11,N=R*C*L;
    double *m,*b,*z,*x0,*x1,*part; uint16_t *pid; int *table;
    CK(cudaMalloc(&m,N*10*8)); CK(cudaMalloc(&b,N*8)); CK(cudaMalloc(&z,N*8)); CK(cudaMalloc(&x0,N*8)); CK(cudaMalloc(&x1,N*8));
    CK(cudaMalloc(&part,8192*8)); CK(cudaMalloc(&pid,N*2)); CK(cudaMalloc(&table,1024*10*4));
    CK(cudaMemset(m,0,N*10*8)); CK(cudaMemset(b,0,N*8)); CK(cudaMemset(x0,0,N*8)); CK(cudaMemset(pid,0,N*2)); CK(cudaMemset(table,0,1024*10*4));
    // z = -2 so that psi > 1 on every row (the common unsaturated case: one division per row)
    { double *h=(double*)malloc(N*8); for(size_t i=0;i<N;++i)h[i]=-2.0; CK(cudaMemcpy(z,h,N*8,cudaMemcpyHostToDevice)); free(h); }
    Big v{}; v.m=m; v.b=b; v.z=z; v.pid=pid; v.table=table; v.part=part; v.N=N; v.hotPid=0;
    const long RC=(long)(R*C);
    for (int c=0;c<10;++c) v.hot[c] = (int)((c==0)? -RC : (c==9)? RC : (c<=3? -(long)C + (c-2) : (c==4? -1 : (c==5? 1 : (long)C + (c-7)))));
    for (int blocks : {1184}) {
        printf("blocks=%d base %.4f ms | +norm %.4f | +pid %.4f | +norm+pid %.4f\n", blocks,
               run<0,0>(v,x0,x1,blocks), run<1,0>(v,x0,x1,blocks), run<0,1>(v,x0,x1,blocks), run<1,1>(v,x0,x1,blocks));
    }
    return 0;
}
