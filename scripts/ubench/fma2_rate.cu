// Micro-benchmark: issue rate of packed fp32x2 vs scalar fp32 arithmetic on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma2_rate fma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){u64 d;asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c));return d;}
__device__ __forceinline__ u64 add2(u64 a, u64 b){u64 d;asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b));return d;}
__device__ __forceinline__ float fma1(float a, float b, float c){float d;asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(d):"f"(a),"f"(b),"f"(c));return d;}
template<int MODE> __global__ void k(float* out, int iters, float s){
  u64 a[8]; float f[8];
  for(int i=0;i<8;++i){a[i]=((u64)__float_as_uint(s+i)<<32)|__float_as_uint(s*i); f[i]=s*i;}
  u64 m=((u64)__float_as_uint(s)<<32)|__float_as_uint(s);
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<8;++i){
      if(MODE==0) a[i]=fma2(a[i],m,a[i]);            // 8 packed FMA
      if(MODE==1) f[i]=fma1(f[i],s,f[i]);            // 8 scalar FMA
      if(MODE==2) a[i]=add2(a[i],m);                 // 8 packed ADD
      if(MODE==3){ a[i]=fma2(a[i],m,a[i]); f[i]=fma1(f[i],s,f[i]); }  // 8 packed + 8 scalar
      if(MODE==4){ a[i]=fma2(a[i],m,a[i]); f[i]=fma1(f[i],s,f[i]); f[i]=fma1(f[i],s,f[i]); } // 8 packed + 16 scalar
    }
  }
  float r=0; for(int i=0;i<8;++i) r+=__uint_as_float((unsigned)a[i])+__uint_as_float((unsigned)(a[i]>>32))+f[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int MODE> void run(const char* name,int packed,int scalar){
  float* out; cudaMalloc(&out,148*8*1024*4);
  int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int w: {4,8,16}) {
    k<MODE><<<148, w*32*4>>>(out, 100, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148, w*32*4>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double winst=(double)iters*8*(packed+scalar)*w*4;      // warp instr per SM
    double cyc=ms*1e-3*1.965e9;
    printf("%-28s warps/SMSP=%2d  ms=%.3f  warp-instr/clk/SMSP=%.3f  fp32 lane-ops/clk/SM=%.1f\n",name,w,ms,winst/cyc/4,(double)iters*8*(packed*64+scalar*32)*w*4/cyc);
  }
  cudaFree(out);
}
int main(){ run<0>("FFMA2 only",1,0); run<1>("FFMA only",0,1); run<2>("FADD2 only",1,0); run<3>("FFMA2 + FFMA 1:1",1,1); run<4>("FFMA2 + FFMA 1:2",1,2); return 0; }
