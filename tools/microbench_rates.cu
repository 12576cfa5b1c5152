// microbench: DFMA, F2F, MUFU, smem atomic throughput on the B200
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters){ double a=threadIdx.x*1e-3,b=1.0000001,c=1e-9; double a2=a+1,a3=a+2,a4=a+3;
  for(int i=0;i<iters;i++){ a=fma(a,b,c); a2=fma(a2,b,c); a3=fma(a3,b,c); a4=fma(a4,b,c);} out[blockIdx.x*blockDim.x+threadIdx.x]=a+a2+a3+a4; }
__global__ void ffma(float* out, int iters){ float a=threadIdx.x*1e-3f,b=1.0000001f,c=1e-9f; float a2=a+1,a3=a+2,a4=a+3;
  for(int i=0;i<iters;i++){ a=fmaf(a,b,c); a2=fmaf(a2,b,c); a3=fmaf(a3,b,c); a4=fmaf(a4,b,c);} out[blockIdx.x*blockDim.x+threadIdx.x]=a+a2+a3+a4; }
__global__ void f2f(double* out, int iters){ double a=threadIdx.x*1e-3+1.0, a2=a+1,a3=a+2,a4=a+3;
  for(int i=0;i<iters;i++){ a=(double)((float)a)+1e-7; a2=(double)((float)a2)+1e-7;a3=(double)((float)a3)+1e-7;a4=(double)((float)a4)+1e-7;} out[blockIdx.x*blockDim.x+threadIdx.x]=a+a2+a3+a4; }
__global__ void mufu(float* out, int iters){ float a=threadIdx.x*1e-3f+1.f,a2=a+1,a3=a+2,a4=a+3;
  for(int i=0;i<iters;i++){ asm("ex2.approx.ftz.f32 %0, %0;":"+f"(a)); asm("lg2.approx.ftz.f32 %0, %0;":"+f"(a)); asm("ex2.approx.ftz.f32 %0, %0;":"+f"(a2)); asm("lg2.approx.ftz.f32 %0, %0;":"+f"(a2));asm("ex2.approx.ftz.f32 %0, %0;":"+f"(a3)); asm("lg2.approx.ftz.f32 %0, %0;":"+f"(a3));asm("ex2.approx.ftz.f32 %0, %0;":"+f"(a4)); asm("lg2.approx.ftz.f32 %0, %0;":"+f"(a4));} out[blockIdx.x*blockDim.x+threadIdx.x]=a+a2+a3+a4; }
__global__ void i2d(double* out, int iters){ int k=threadIdx.x; double s=0,s2=0;
  for(int i=0;i<iters;i++){ s+= (double)(k+i); s2 += (double)(k-i); k = (int)(s*1e-3);} out[blockIdx.x*blockDim.x+threadIdx.x]=s+s2; }
__global__ void ffma2k(float* out, int iters){ float2 a=make_float2(threadIdx.x*1e-3f,1.f),b=make_float2(1.0000001f,0.999999f),c=make_float2(1e-9f,1e-8f); float2 a2=a,a3=a,a4=a; a2.x+=1;a3.x+=2;a4.x+=3;
  for(int i=0;i<iters;i++){ a=__ffma2_rn(a,b,c); a2=__ffma2_rn(a2,b,c); a3=__ffma2_rn(a3,b,c); a4=__ffma2_rn(a4,b,c);} out[blockIdx.x*blockDim.x+threadIdx.x]=a.x+a2.x+a3.x+a4.x+a.y+a2.y+a3.y+a4.y; }
__global__ void ffma2lop(float* out, int iters){ float2 a=make_float2(threadIdx.x*1e-3f,1.f),b=make_float2(1.0000001f,0.999999f),c=make_float2(1e-9f,1e-8f); float2 a2=a,a3=a,a4=a; a2.x+=1;a3.x+=2;a4.x+=3; unsigned m=0xffffffffu;
  for(int i=0;i<iters;i++){ a=__ffma2_rn(a,b,c); a2=__ffma2_rn(a2,b,c); a3=__ffma2_rn(a3,b,c); a4=__ffma2_rn(a4,b,c); m &= __float_as_uint(a.x)&__float_as_uint(a2.y); } out[blockIdx.x*blockDim.x+threadIdx.x]=a.x+a2.x+a3.x+a4.x+a.y+a2.y+a3.y+a4.y+(float)m; }
__global__ void dsetp(double* out, int iters){ double a=threadIdx.x*1e-3,b=1.0000001,c=1e-9; double m1=1e300,m2=1e300,m3=1e300,m4=1e300; double a2=a+1,a3=a+2,a4=a+3;
  for(int i=0;i<iters;i++){ a=fma(a,b,c); a2=fma(a2,b,c); a3=fma(a3,b,c); a4=fma(a4,b,c); m1 = a<m1?a:m1; m2=a2<m2?a2:m2; m3=a3<m3?a3:m3; m4=a4<m4?a4:m4;} out[blockIdx.x*blockDim.x+threadIdx.x]=m1+m2+m3+m4; }
template<int MODE> __global__ void satom(unsigned* out, int iters){ __shared__ unsigned h[2048]; for(int i=threadIdx.x;i<2048;i+=blockDim.x) h[i]=0; __syncthreads();
  unsigned x = threadIdx.x*2654435761u + blockIdx.x;
  for(int i=0;i<iters;i++){ x = x*1664525u+1013904223u; unsigned idx;
    if(MODE==0) idx = (x>>8)&2047;            // random spread
    else if(MODE==1) idx = 5;                 // all same address
    else if(MODE==2) idx = ((i>>4)&255)*8 + (threadIdx.x&7);  // same bin, 8 copies
    else idx = (((i>>4)+((x>>28)&3))&255)*8 + (threadIdx.x&7); // 4 neighbouring bins x 8 copies
    atomicAdd(&h[idx],1u);} __syncthreads(); if(threadIdx.x==0) out[blockIdx.x]=h[5]+h[7]; }
template<class F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a);cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }
int main(){ int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount,0); void* buf; cudaMalloc(&buf, 1<<26);
  int blocks = sm*8, thr=256, iters=20000;
  float ms;
  ms=timeit([&]{dfma<<<blocks,thr>>>((double*)buf,iters);}); printf("DFMA  %.2f T inst/s (%.1f TFLOP/s)\n", 4.0*iters*blocks*thr/ms/1e9, 8.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{ffma<<<blocks,thr>>>((float*)buf,iters);}); printf("FFMA  %.2f T inst/s\n", 4.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{ffma2k<<<blocks,thr>>>((float*)buf,iters);}); printf("FFMA2 %.2f T inst/s (2 fma each)\n", 4.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{ffma2lop<<<blocks,thr>>>((float*)buf,iters);}); printf("FFMA2 x4 + LOP3 %.2f T inst/s (FFMA2 only counted)\n", 4.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{dsetp<<<blocks,thr>>>((double*)buf,iters);}); printf("DFMA + fp64 min (compare-select) %.2f T pairs/s\n", 4.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{f2f<<<blocks,thr>>>((double*)buf,iters);}); printf("F2F pair+DADD %.2f T triples/s\n", 4.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{mufu<<<blocks,thr>>>((float*)buf,iters);}); printf("MUFU  %.2f T inst/s\n", 8.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{i2d<<<blocks,thr>>>((double*)buf,iters);}); printf("I2D/D2I loop %.2f T iter/s\n", 1.0*iters*blocks*thr/ms/1e9);
  iters=4000;
  ms=timeit([&]{satom<0><<<blocks,thr>>>((unsigned*)buf,iters);}); printf("ATOMS random    %.3f T atom/s\n", 1.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{satom<1><<<blocks,thr>>>((unsigned*)buf,iters);}); printf("ATOMS same addr %.3f T atom/s\n", 1.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{satom<2><<<blocks,thr>>>((unsigned*)buf,iters);}); printf("ATOMS 1bin x8   %.3f T atom/s\n", 1.0*iters*blocks*thr/ms/1e9);
  ms=timeit([&]{satom<3><<<blocks,thr>>>((unsigned*)buf,iters);}); printf("ATOMS 4bin x8   %.3f T atom/s\n", 1.0*iters*blocks*thr/ms/1e9);
  return 0; }
