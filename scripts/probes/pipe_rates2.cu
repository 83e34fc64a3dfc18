#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int ILP=8, ITERS=4096;
__device__ __forceinline__ __half2 H(uint32_t u){return *reinterpret_cast<__half2*>(&u);}
__device__ __forceinline__ uint32_t Ub(__half2 h){return *reinterpret_cast<uint32_t*>(&h);}
template<int W> __global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b){
  uint32_t x[ILP], y[ILP];
  for(int i=0;i<ILP;i++){ x[i]=threadIdx.x*65537u+i; y[i]=x[i]^0x1234;}
  for(int it=0;it<ITERS;it++){
    __half2 ha=H(a), hb=H(b);
    #pragma unroll
    for(int i=0;i<ILP;i++){
      if(W==0){ x[i]=Ub(__hlt2(H(x[i]),ha)); }                       // HSET2
      if(W==1){ x[i]=Ub(__hlt2(H(x[i]),ha)); y[i]=Ub(__hadd2(H(y[i]),hb)); }   // HSET2 + HADD2
      if(W==2){ x[i]=Ub(__hmax2(H(x[i]),ha)); y[i]=Ub(__hadd2(H(y[i]),hb)); }  // HMNMX2 + HADD2
      if(W==3){ x[i]=Ub(__hmax2(H(x[i]),ha)); y[i]=(y[i]>>1)|(x[i]&0x80008000u);} // HMNMX2 + SHF + LOP3
      if(W==4){ x[i]=Ub(__hfma2_relu(H(x[i]),ha,hb)); }               // HFMA2.RELU
      if(W==5){ x[i]=Ub(__hmax2(H(x[i]),ha)); y[i]=__funnelshift_l(a,y[i],1);} // HMNMX2 + SHF
      if(W==6){ y[i]=(y[i]>>1)|(a&0x80008000u);}                      // SHF+LOP3
      if(W==7){ y[i]=y[i]*2+(x[i]&1); x[i]+=a;}                       // IMAD-ish accumulate
    }
    a^=it;
  }
  uint32_t s=0; for(int i=0;i<ILP;i++) s+=x[i]+y[i];
  if(s==0x12345) out[0]=s;
}
template<int W> double run(uint32_t* d){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best=1e9;
  for(int r=0;r<4;r++){ cudaEventRecord(e0); k<W><<<148*8,256>>>(d,0x3c003c00u,0x40004000u); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r&&ms<best)best=ms;}
  return 148.0*8*256*ITERS*ILP/(best*1e-3)/1e12;
}
int main(){ uint32_t* d; cudaMalloc(&d,1024);
 printf("HSET2 %.2f T iter/s\nHSET2+HADD2 %.2f\nHMNMX2+HADD2 %.2f\nHMNMX2+SHF+LOP3 %.2f\nHFMA2.RELU %.2f\nHMNMX2+SHF %.2f\nSHF+LOP3 %.2f\nIMAD acc %.2f\n", run<0>(d),run<1>(d),run<2>(d),run<3>(d),run<4>(d),run<5>(d),run<6>(d),run<7>(d)); return 0; }
