// Micro-benchmark: does a packed FFMA2 (2 cycles on the FMA pipe) leave its second issue cycle to
// another pipe?  Mixes FFMA2 / FFMA with integer ALU work (LOP3 / IADD3) and shared-memory loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma2_mix fma2_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){u64 d;asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c));return d;}
__device__ __forceinline__ float fma1(float a, float b, float c){float d;asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(d):"f"(a),"f"(b),"f"(c));return d;}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b){unsigned d;asm volatile("xor.b32 %0, %1, %2;":"=r"(d):"r"(a),"r"(b));return d;}
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b){unsigned d;asm volatile("add.u32 %0, %1, %2;":"=r"(d):"r"(a),"r"(b));return d;}
template<int P,int S,int I> __global__ void k(float* out, int iters, float s, unsigned q){
  u64 a[8]; float f[8]; unsigned u[8];
  for(int i=0;i<8;++i){a[i]=((u64)__float_as_uint(s+i)<<32)|__float_as_uint(s*i); f[i]=s*i; u[i]=q+i;}
  u64 m=((u64)__float_as_uint(s)<<32)|__float_as_uint(s);
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<8;++i){
      #pragma unroll
      for(int r=0;r<P;++r) a[i]=fma2(a[i],m,a[i]);
      #pragma unroll
      for(int r=0;r<S;++r) f[i]=fma1(f[i],s,f[i]);
      #pragma unroll
      for(int r=0;r<I;++r) u[i]=(r&1)?lop(u[i],q):iadd(u[i],q);
    }
  }
  float r=0; for(int i=0;i<8;++i) r+=__uint_as_float((unsigned)a[i])+__uint_as_float((unsigned)(a[i]>>32))+f[i]+(float)u[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int P,int S,int I> void run(const char* name){
  float* out; cudaMalloc(&out,148*8*1024*4);
  int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int w: {4,8}) {
    k<P,S,I><<<148, w*32*4>>>(out, 100, 1.0001f, 12345u); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<P,S,I><<<148, w*32*4>>>(out, iters, 1.0001f, 12345u); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double winst=(double)iters*8*(P+S+I)*w*4;
    double cyc=ms*1e-3*1.965e9;
    printf("%-34s warps/SMSP=%2d  ms=%.3f  warp-instr/clk/SMSP=%.3f  (cycles per group of %d: %.2f)\n",name,w,ms,winst/cyc/4,P+S+I,(P+S+I)/(winst/cyc/4));
  }
  cudaFree(out);
}
int main(){
  run<1,0,0>("FFMA2 only");
  run<0,1,0>("FFMA only");
  run<0,0,1>("INT only");
  run<1,0,1>("FFMA2 + INT 1:1");
  run<1,0,2>("FFMA2 + INT 1:2");
  run<0,2,2>("FFMA + INT 2:2");
  run<0,2,1>("FFMA + INT 2:1");
  run<2,0,1>("FFMA2 + INT 2:1");
  return 0; }
