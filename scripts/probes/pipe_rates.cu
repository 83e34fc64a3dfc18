#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int ILP=8, ITERS=4096;
template<int W> __global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b){
  uint32_t x[ILP];
  for(int i=0;i<ILP;i++) x[i]=threadIdx.x*65537u+i;
  __half2 ha=*reinterpret_cast<__half2*>(&a), hb=*reinterpret_cast<__half2*>(&b);
  for(int it=0;it<ITERS;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++){
      __half2 h=*reinterpret_cast<__half2*>(&x[i]);
      if(W==0) h=__hadd2(h,ha);
      if(W==1) h=__hmax2(h,hb);
      if(W==2) h=__hfma2(h,ha,hb);
      if(W==3) { x[i]=(x[i]>>1)|(a&0x80008000u); continue; }
      if(W==4) { x[i]=__funnelshift_l(a,x[i],1); continue; }
      if(W==5) { x[i]=max((int)x[i],(int)b)+ (int)a; continue; }
      x[i]=*reinterpret_cast<uint32_t*>(&h);
    }
    a^=it;
    ha=*reinterpret_cast<__half2*>(&a);
  }
  uint32_t s=0; for(int i=0;i<ILP;i++) s+=x[i];
  if(s==0x12345) out[0]=s;
}
template<int W> double run(uint32_t* d){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best=1e9;
  for(int r=0;r<4;r++){ cudaEventRecord(e0); k<W><<<148*8,256>>>(d,0x3c003c00u,0x40004000u); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r&&ms<best)best=ms;}
  return 148.0*8*256*ITERS*ILP/(best*1e-3)/1e12;
}
int main(){ uint32_t* d; cudaMalloc(&d,1024);
 printf("HADD2 %.2f T/s\nHMNMX2 %.2f\nHFMA2 %.2f\nSHR+LOP3(2 instr) %.2f\nSHF.L funnel %.2f\nmax+add %.2f\n", run<0>(d),run<1>(d),run<2>(d),run<3>(d),run<4>(d),run<5>(d)); return 0; }
