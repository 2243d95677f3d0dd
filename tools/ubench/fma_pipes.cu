// Microbenchmark: per-SM throughput of FFMA, FFMA2, FMNMX3 and mixes (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b){u64 r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ u64 ffma2(u64 a,u64 b,u64 c){u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
__device__ __forceinline__ float ffma(float a,float b,float c){float d; asm volatile("fma.rn.f32 %0,%1,%2,%3;":"=f"(d):"f"(a),"f"(b),"f"(c)); return d;}
__device__ __forceinline__ float min3(float a,float b,float c){float d; asm volatile("min.f32 %0,%1,%2,%3;":"=f"(d):"f"(a),"f"(b),"f"(c)); return d;}
__device__ __forceinline__ unsigned lop(unsigned a,unsigned b){unsigned d; asm volatile("and.b32 %0,%1,%2;":"=r"(d):"r"(a),"r"(b)); return d;}

template<int MODE> __global__ void k(float* out, int iters, float s){
  float a[8]; u64 p[8]; unsigned q[4];
  for(int i=0;i<8;i++){a[i]=threadIdx.x*0.001f+i; p[i]=pack2(a[i],a[i]+1);} for(int i=0;i<4;i++) q[i]=threadIdx.x+i;
  u64 sp=pack2(s,s);
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int r=0;r<8;r++){
      if(MODE==0){ for(int i=0;i<8;i++) a[i]=ffma(a[i],s,a[(i+1)&7]); }          // 8 FFMA
      if(MODE==1){ for(int i=0;i<8;i++) p[i]=ffma2(p[i],sp,p[(i+1)&7]); }        // 8 FFMA2
      if(MODE==2){ for(int i=0;i<4;i++){ p[i]=ffma2(p[i],sp,p[(i+1)&3]); a[i]=ffma(a[i],s,a[(i+1)&3]); a[i+4]=ffma(a[i+4],s,a[4+((i+1)&3)]);} } // 4 FFMA2 + 8 FFMA
      if(MODE==3){ for(int i=0;i<8;i++){ p[i]=ffma2(p[i],sp,p[(i+1)&7]);} for(int i=0;i<4;i++) a[i]=min3(a[i],a[(i+1)&3],a[(i+2)&3]); } // 8 FFMA2 + 4 FMNMX3
      if(MODE==4){ for(int i=0;i<8;i++){ a[i]=ffma(a[i],s,a[(i+1)&7]);} for(int i=0;i<4;i++) q[i]=lop(q[i],q[(i+1)&3]); } // 8 FFMA + 4 LOP
      if(MODE==5){ for(int i=0;i<8;i++){ p[i]=ffma2(p[i],sp,p[(i+1)&7]);} for(int i=0;i<8;i++) q[i&3]=lop(q[i&3],q[(i+1)&3]); } // 8 FFMA2 + 8 LOP
      if(MODE==6){ for(int i=0;i<4;i++) a[i]=min3(a[i],a[(i+1)&3],a[(i+2)&3]); for(int i=0;i<4;i++) q[i]=lop(q[i],q[(i+1)&3]); } // 4 FMNMX3+4 LOP
    }
  }
  float acc=0; for(int i=0;i<8;i++){acc+=a[i]; float lo,hi; asm("mov.b64 {%0,%1},%2;":"=f"(lo),"=f"(hi):"l"(p[i])); acc+=lo+hi;} for(int i=0;i<4;i++) acc+=q[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
template<int MODE> void run(const char* name, double fma_per_iter, double inst_per_iter, int warps){
  float* out; cudaMalloc(&out, 148*1024*4*4);
  int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, warps*32>>>(out, 100, 1.0001f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148, warps*32>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double clk=1.965e9*ms*1e-3; // cycles (assuming 1965 MHz)
  double warp_inst=(double)iters*8*inst_per_iter*warps; // per SM
  double fmas=(double)iters*8*fma_per_iter*warps*32;
  printf("%-28s warps/SM=%2d  %.3f ms  warp-inst/clk/SM=%.2f  FMA/clk/SM=%.1f\n", name, warps, ms, warp_inst/clk, fmas/clk);
  cudaFree(out);
}
int main(){
  for(int w: {4,8,16,32}){
    run<0>("FFMA x8", 8, 8, w);
    run<1>("FFMA2 x8", 16, 8, w);
    run<2>("FFMA2 x4 + FFMA x8", 16, 12, w);
    run<3>("FFMA2 x8 + FMNMX3 x4", 16, 12, w);
    run<4>("FFMA x8 + LOP x4", 8, 12, w);
    run<5>("FFMA2 x8 + LOP x8", 16, 16, w);
    run<6>("FMNMX3 x4 + LOP x4", 0, 8, w);
  }
  return 0;
}
