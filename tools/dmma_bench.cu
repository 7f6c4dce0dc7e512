// Micro-benchmark: FP64 pipe throughput on sm_100a (DMMA.8x8x4 vs DFMA), used to pick
// the consumer side of the first-contraction kernel and as a sanity bound for the roofline.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)

template<int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters, double av, double bv){
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  double a=av+threadIdx.x*1e-9, b=bv;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++){
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
    }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters, double av, double bv){
  double c[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  double a=av+threadIdx.x*1e-9, b=bv;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(a,c[i],b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}


// DMMA and DFMA interleaved: do they share one FP64 pipe?  NM DMMA + NF DFMA per iteration, independent accumulators.
template<int NM, int NF>
__global__ void __launch_bounds__(1024) mixed_kernel(double* out, int iters, double av, double bv){
  double c[NM][2]; double f[NF > 0 ? NF : 1];
  #pragma unroll
  for(int i=0;i<NM;i++){c[i][0]=0;c[i][1]=0;}
  #pragma unroll
  for(int i=0;i<NF;i++) f[i]=i;
  double a=av+threadIdx.x*1e-9, b=bv;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NM;i++){
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
      #pragma unroll
      for(int j=0;j<NF/NM;j++) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(f[i*(NF/NM)+j]) : "d"(a),"d"(b));
    }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NM;i++) s+=c[i][0]+c[i][1];
  #pragma unroll
  for(int i=0;i<NF;i++) s+=f[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NM,int NF>
void run_mixed(double* out,int sms,cudaEvent_t e0,cudaEvent_t e1){
  const int iters=20000, threads=128, bps=2;
  mixed_kernel<NM,NF><<<sms*bps,threads>>>(out,100,0.999,1e-3);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  mixed_kernel<NM,NF><<<sms*bps,threads>>>(out,iters,0.999,1e-3);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
  double fm=2.0*8*8*4*NM*(double)iters*(threads/32)*sms*bps, ff=2.0*NF*(double)iters*threads*sms*bps;
  printf("mixed NM=%d DMMA + NF=%d DFMA per iter: %.3f ms  DMMA-only-equivalent %.2f TF/s, total %.2f TF/s\n",NM,NF,ms,fm/ms*1e-9,(fm+ff)/ms*1e-9);
}

int main(){
  int dev=0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int sms=p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n",p.name,sms,p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*4*1024));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int iters=20000;
  for(int threads : {128,256,512,1024}){
    for(int bps : {1,2}){
      if(threads*bps>2048) continue;
      // DMMA with 8 independent accumulators
      dmma_kernel<8><<<sms*bps,threads>>>(out,100,1.0,1.0);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      dmma_kernel<8><<<sms*bps,threads>>>(out,iters,1.0,1e-3);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
      double flops=2.0*8*8*4*8.0*iters*(threads/32)*sms*bps;
      printf("DMMA.8x8x4 nacc=8 threads=%4d blocks/SM=%d : %.2f TFLOP/s (%.3f ms)\n",threads,bps,flops/ms*1e-9,ms);
      dmma_kernel<16><<<sms*bps,threads>>>(out,iters,1.0,1e-3);
      CK(cudaEventRecord(e0));
      dmma_kernel<16><<<sms*bps,threads>>>(out,iters,1.0,1e-3);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms,e0,e1));
      flops=2.0*8*8*4*16.0*iters*(threads/32)*sms*bps;
      printf("DMMA.8x8x4 nacc=16 threads=%4d blocks/SM=%d : %.2f TFLOP/s (%.3f ms)\n",threads,bps,flops/ms*1e-9,ms);
      dfma_kernel<16><<<sms*bps,threads>>>(out,iters,1.0,1e-3);
      CK(cudaEventRecord(e0));
      dfma_kernel<16><<<sms*bps,threads>>>(out,iters,0.999,1e-3);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms,e0,e1));
      flops=2.0*16.0*iters*threads*sms*bps;
      printf("DFMA       nacc=16 threads=%4d blocks/SM=%d : %.2f TFLOP/s (%.3f ms)\n",threads,bps,flops/ms*1e-9,ms);
    }
  }
  run_mixed<8,0>(out,sms,e0,e1); run_mixed<8,8>(out,sms,e0,e1); run_mixed<8,16>(out,sms,e0,e1); run_mixed<8,32>(out,sms,e0,e1); run_mixed<8,64>(out,sms,e0,e1);
  // latency of a dependent DMMA chain (1 warp)
  CK(cudaEventRecord(e0));
  dmma_kernel<1><<<1,32>>>(out,200000,1.0,1e-3);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
  printf("dependent DMMA chain: %.2f ns per DMMA\n", ms*1e6/200000);
  return 0;
}
