// microbenchmark: peer store bandwidth (8B vs 16B per thread, with/without per-block sys fence), flag ping latency
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
__global__ void st8(double* dst, const double* src, long n, int fence){
  long i=(long)blockIdx.x*blockDim.x*4+threadIdx.x;
  #pragma unroll
  for(int j=0;j<4;++j){ long t=i+j*256; if(t<n) dst[t]=src[t]; }
  if(fence){ __syncthreads(); if(threadIdx.x==0) __threadfence_system(); }
}
__global__ void st16(double2* dst, const double2* src, long n2, int fence){
  long i=(long)blockIdx.x*blockDim.x*2+threadIdx.x;
  #pragma unroll
  for(int j=0;j<2;++j){ long t=i+j*256; if(t<n2) dst[t]=src[t]; }
  if(fence){ __syncthreads(); if(threadIdx.x==0) __threadfence_system(); }
}
__global__ void st8_persist(double* dst, const double* src, long n){
  for(long t=(long)blockIdx.x*blockDim.x+threadIdx.x;t<n;t+=(long)gridDim.x*blockDim.x) dst[t]=src[t];
  __syncthreads(); if(threadIdx.x==0) __threadfence_system();
}
__global__ void empty_fence(unsigned long long* flag, unsigned long long v){ __threadfence_system(); asm volatile("st.release.sys.global.u64 [%0], %1;"::"l"(flag),"l"(v):"memory"); }
int main(){
  int nd=0; CK(cudaGetDeviceCount(&nd)); if(nd<2){printf("need 2 gpus\n");return 0;}
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1,0));
  CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0,0));
  for(long mb: {1L,3L,8L,32L}){
    long n=mb*1024*1024/8;
    double *src,*dloc,*drem;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&drem,n*8));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&src,n*8)); CK(cudaMalloc(&dloc,n*8)); CK(cudaMemset(src,0,n*8));
    cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    auto run=[&](const char* name, auto fn){ float best=1e9; for(int r=0;r<20;++r){ cudaEventRecord(a); fn(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;} printf("%4ld MB %-28s %8.2f us  %8.1f GB/s\n",mb,name,best*1e3,n*8/(best*1e-3)/1e9); };
    int nb4=(n+1023)/1024, nb2=(n/2+511)/512;
    run("local st8",[&]{st8<<<nb4,256>>>(dloc,src,n,0);});
    run("peer st8",[&]{st8<<<nb4,256>>>(drem,src,n,0);});
    run("peer st8 + block fence",[&]{st8<<<nb4,256>>>(drem,src,n,1);});
    run("peer st16",[&]{st16<<<nb2,256>>>((double2*)drem,(const double2*)src,n/2,0);});
    run("peer st16 + block fence",[&]{st16<<<nb2,256>>>((double2*)drem,(const double2*)src,n/2,1);});
    run("peer st8 persistent 148x4",[&]{st8_persist<<<148*4,256>>>(drem,src,n);});
    run("peer st8 persistent 148x8",[&]{st8_persist<<<148*8,256>>>(drem,src,n);});
    run("peer memcpy",[&]{cudaMemcpyPeerAsync(drem,1,src,0,n*8,0);});
    cudaFree(src); cudaFree(dloc); cudaSetDevice(1); cudaFree(drem); cudaSetDevice(0);
  }
  // launch + fence + remote flag store
  unsigned long long* flag; CK(cudaSetDevice(1)); CK(cudaMalloc(&flag,8)); CK(cudaSetDevice(0));
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best=1e9; for(int r=0;r<50;++r){ cudaEventRecord(a); empty_fence<<<1,32>>>(flag,r); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; }
  printf("empty kernel with sys fence + remote flag store: %.2f us\n",best*1e3);
  return 0;
}
