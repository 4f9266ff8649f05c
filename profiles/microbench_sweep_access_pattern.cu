// experiment: how fast can the sweep's memory pattern stream on this GPU? (not part of the product)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__global__ void __launch_bounds__(256) k_stream(const double* __restrict__ m, const double* __restrict__ b, const double* __restrict__ z,
    const double* __restrict__ xin, double* __restrict__ xout, size_t N, int gather, int rowsPerLayer, int cols)
{
    for (size_t i = blockIdx.x*256ull+threadIdx.x; i < N; i += (size_t)gridDim.x*256) {
        double acc = __ldcs(b+i);
        #pragma unroll
        for (int c=0;c<10;++c) {
            double a = __ldcs(m + (size_t)c*N + i);
            double xv = 1.0;
            if (gather) {
                long off = (c==0)? -(long)rowsPerLayer : (c==9)? (long)rowsPerLayer : (c<=3? -(long)cols + (c-2) : (c==4? -1 : (c==5? 1 : (long)cols + (c-7))));
                long j = (long)i + off; if (j<0) j=i; if (j>=(long)N) j=i;
                xv = xin[j];
            }
            acc -= a*xv;
        }
        double zz = __ldcs(z+i);
        double xo = xin[i];
        xout[i] = acc + zz*1e-30 + xo*1e-30;
    }
}
int main(){
    size_t R=1024,C=1024,L=11,N=R*C*L;
    double *m,*b,*z,*x0,*x1;
    CK(cudaMalloc(&m,N*10*8)); CK(cudaMalloc(&b,N*8)); CK(cudaMalloc(&z,N*8)); CK(cudaMalloc(&x0,N*8)); CK(cudaMalloc(&x1,N*8));
    CK(cudaMemset(m,0,N*10*8)); CK(cudaMemset(b,0,N*8)); CK(cudaMemset(z,0,N*8)); CK(cudaMemset(x0,0,N*8));
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int gather=0; gather<2; ++gather) for (int blocks : {1184, 2368, 148*16, 148*32}) {
        for (int w=0;w<3;++w) k_stream<<<blocks,256>>>(m,b,z,x0,x1,N,gather,(int)(R*C),(int)C);
        cudaEventRecord(e0);
        for (int it=0; it<20; ++it) { k_stream<<<blocks,256>>>(m,b,z,(it&1)?x1:x0,(it&1)?x0:x1,N,gather,(int)(R*C),(int)C); }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms,e0,e1); ms/=20;
        printf("gather=%d blocks=%d  %.4f ms  real bytes %.2f GB -> %.0f GB/s\n", gather, blocks, ms, N*(80+8+8+8+8)/1e9, N*112.0/ms/1e6);
    }
    return 0;
}
