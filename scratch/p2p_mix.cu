// microbenchmark 2: does remote-store backpressure block local memory work on the same SMs?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
// each block handles 1024 doubles; kind(block) decides remote or local destination
// mode 0: all local, 1: all remote, 2: interleave (b%3!=0 remote => 2/3 remote), 3: remote first 2/3 then local, 4: local first
__global__ void mix(double* dloc, double* drem, const double* src, long nblk, int mode){
  long b=blockIdx.x; bool remote;
  if(mode==0) remote=false; else if(mode==1) remote=true; else if(mode==2) remote=(b%3)!=0; else if(mode==3) remote= b < nblk*2/3; else remote = b >= nblk/3;
  double* d = remote? drem: dloc;
  long i=b*1024+threadIdx.x;
  double v[4];
  #pragma unroll
  for(int j=0;j<4;++j) v[j]=src[i+j*256]+1.0;
  #pragma unroll
  for(int j=0;j<4;++j) d[i+j*256]=v[j];
}
// TMA bulk store variant: stage 1024 doubles in smem, one thread issues cp.async.bulk smem->global
__global__ void mix_bulk(double* dloc, double* drem, const double* src, long nblk, int mode){
  __shared__ __align__(128) double sm[1024];
  long b=blockIdx.x; bool remote;
  if(mode==0) remote=false; else if(mode==1) remote=true; else if(mode==2) remote=(b%3)!=0; else if(mode==3) remote= b < nblk*2/3; else remote = b >= nblk/3;
  long i=b*1024+threadIdx.x;
  #pragma unroll
  for(int j=0;j<4;++j) sm[threadIdx.x+j*256]=src[i+j*256]+1.0;
  if(!remote){
    #pragma unroll
    for(int j=0;j<4;++j) dloc[i+j*256]=sm[threadIdx.x+j*256];
    return;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if(threadIdx.x==0){
    unsigned s=(unsigned)__cvta_generic_to_shared(sm);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(drem+b*1024), "r"(s), "r"(8192) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
int main(){
  int nd=0; CK(cudaGetDeviceCount(&nd)); if(nd<2){printf("need 2 gpus\n");return 0;}
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1,0));
  long nblk=1560; long n=nblk*1024;   // 12.8 MB total like the N=2 push
  double *src,*dloc,*drem;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&drem,n*8));
  CK(cudaSetDevice(0)); CK(cudaMalloc(&src,n*8)); CK(cudaMalloc(&dloc,n*8)); CK(cudaMemset(src,0,n*8));
  double* flush; CK(cudaMalloc(&flush,256<<20));
  cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const char* names[5]={"all local","all remote","interleaved 2/3 remote","remote-first 2/3","local-first 1/3"};
  for(int bulk=0;bulk<2;++bulk) for(int mode=0;mode<5;++mode){
    float best=1e9,sum=0; int reps=20;
    for(int r=0;r<reps;++r){ cudaMemsetAsync(flush,0,256<<20); cudaEventRecord(a); if(bulk) mix_bulk<<<nblk,256>>>(dloc,drem,src,nblk,mode); else mix<<<nblk,256>>>(dloc,drem,src,nblk,mode); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; sum+=ms;}
    printf("%-6s %-26s best %7.2f us  mean %7.2f us\n",bulk?"bulk":"st8",names[mode],best*1e3,sum/reps*1e3);
  }
  CK(cudaGetLastError());
  return 0;
}
