// tools/ubench.cu -- instruction-throughput microbenchmarks for sm_100a (tuning aid, not product).
// Each kernel runs ITER trips of an unrolled body of independent instruction chains on every SM at
// full occupancy; we report warp-instructions per cycle per SM (4 = one per SMSP per cycle).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long f2;
#define ITER 1024
#define REP 8
#define NCH 8

__device__ __forceinline__ f2 pk(float a, float b){ f2 r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r; }

template<int MODE> __global__ void __launch_bounds__(512) k(float* out, float seed, long long* cyc) {
  float a[NCH]; f2 p[NCH]; uint32_t u[NCH];
  for (int i=0;i<NCH;++i){ a[i]=seed+i+threadIdx.x; p[i]=pk(a[i],a[i]+1.f); u[i]=threadIdx.x*2654435761u+i; }
  const float c1=seed*0.5f, c2=seed*0.25f; const f2 q1=pk(c1,c1), q2=pk(c2,c2);
  long long t0=clock64();
  #pragma unroll 1
  for (int it=0; it<ITER; ++it) {
    #pragma unroll
    for (int rep=0; rep<REP; ++rep)
    #pragma unroll
    for (int i=0;i<NCH;++i) {
      if (MODE==0) a[i]=fmaf(a[i],c1,c2);                                           // FFMA
      if (MODE==1) asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p[i]):"l"(q1),"l"(q2)); // FFMA2
      if (MODE==2) u[i]=__byte_perm(u[i],0x4B000000u,0x7541);                       // PRMT
      if (MODE==3) asm volatile("lop3.b32 %0,%0,%1,%2,0x78;":"+r"(u[i]):"r"(0x80808080u),"r"(u[(i+1)%NCH])); // LOP3
      if (MODE==4) a[i]=fminf(fabsf(a[i]),c1);                                      // FMNMX
      if (MODE==5) { a[i]=fmaf(a[i],c1,c2); u[i]=__byte_perm(u[i],0x4B000000u,0x7541);} // FFMA+PRMT
      if (MODE==6) { asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p[i]):"l"(q1),"l"(q2)); u[i]=__byte_perm(u[i],0x4B000000u,0x7541);} // FFMA2+PRMT
      if (MODE==7) { asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p[i]):"l"(q1),"l"(q2)); a[i]=fmaf(a[i],c1,c2);} // FFMA2+FFMA
      if (MODE==8) asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(q1));       // FADD2
      if (MODE==9) a[i]=a[i]+c1;                                                    // FADD
      if (MODE==10) asm volatile("add.u32 %0,%0,%1;":"+r"(u[i]):"r"(u[(i+3)%NCH]));                                       // IADD3
      if (MODE==11) asm volatile("rcp.approx.ftz.f32 %0,%0;":"+f"(a[i]));           // MUFU
      if (MODE==12) { asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p[i]):"l"(q1),"l"(q2)); u[i]=__byte_perm(u[i],0x4B000000u,0x7541); u[(i+1)%NCH]^=u[i]>>3; } // FFMA2 + 2 ALU (PRMT, LOP/SHF)
      if (MODE==13) { a[i]=fmaf(a[i],c1,c2); a[i]=fmaf(a[i],c2,c1); u[i]=__byte_perm(u[i],0x4B000000u,0x7541);} // 2 FFMA + PRMT
      if (MODE==14) { a[i]=fmaf(a[i],c1,c2); asm volatile("lop3.b32 %0,%0,%1,%2,0x78;":"+r"(u[i]):"r"(0x80808080u),"r"(u[(i+1)%NCH])); u[i]=__byte_perm(u[i],0x4B000000u,0x7541);} // FFMA + 2 ALU
      if (MODE==15) asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(q1));      // FMUL2
      if (MODE==16) { float t; asm volatile("{.reg .b8 b0,b1,b2,b3; mov.b32 {b0,b1,b2,b3}, %1; cvt.rn.f32.s8 %0, b1;}":"=f"(t):"r"(u[i])); a[i]=t; u[i]=__float_as_uint(t); } // I2F.S8 chain
      if (MODE==17) { float t; asm volatile("{.reg .b8 b0,b1,b2,b3; mov.b32 {b0,b1,b2,b3}, %1; cvt.rn.f32.s8 %0, b1;}":"=f"(t):"r"(u[i])); a[i]+=t; } // I2F.S8 (indep) + FADD
      if (MODE==18) { float t; asm volatile("{.reg .b8 b0,b1,b2,b3; mov.b32 {b0,b1,b2,b3}, %1; cvt.rn.f32.s8 %0, b2;}":"=f"(t):"r"(u[i])); a[i]+=t; u[i]=__byte_perm(u[i],0x4B000000u,0x7541);} // I2F + FADD + PRMT
      if (MODE==19) { float t; asm volatile("{.reg .b8 b0,b1,b2,b3; mov.b32 {b0,b1,b2,b3}, %1; cvt.rn.f32.s8 %0, b2;}":"=f"(t):"r"(u[i])); a[i]=fmaf(a[i],c1,t); a[i]=fmaf(a[i],c2,c1); a[i]=fmaf(a[i],c1,c2); } // I2F + 3 FFMA
      if (MODE==20) { a[i]=fmaf(a[i],c1,c2); a[i]=fmaf(a[i],c2,c1); a[i]=fmaf(a[i],c1,c2); a[i]=fmaf(a[i],c2,c1); u[i]=__byte_perm(u[i],0x4B000000u,0x7541);} // 4 FFMA + PRMT
      if (MODE==21) { asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p[i]):"l"(q1),"l"(q2)); u[i]=__byte_perm(u[i],0x4B000000u,0x7541); asm volatile("lop3.b32 %0,%0,%1,%2,0x78;":"+r"(u[(i+4)%NCH]):"r"(0x80808080u),"r"(u[(i+1)%NCH])); } // FFMA2 + PRMT + LOP3 (indep)
    }
  }
  long long t1=clock64();
  float s=0; uint32_t x=0; for (int i=0;i<NCH;++i){ s+=a[i]; float lo,hi; asm("mov.b64 {%0,%1},%2;":"=f"(lo),"=f"(hi):"l"(p[i])); s+=lo+hi; x^=u[i]; }
  if (s==123.456f || x==0x12345u) out[0]=s;
  if (threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}

template<int MODE> void run(const char* name, int instr_per_chain) {
  int dev=0, sms=0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out; long long* cyc; cudaMalloc(&out,4); cudaMalloc(&cyc, sizeof(long long)*sms*2);
  const int blocks=sms;  // 1 x 512 threads = 16 warps per SM (4 per SMSP), NCH independent chains each
  k<MODE><<<blocks,512>>>(out,1.0f,cyc); cudaDeviceSynchronize();
  k<MODE><<<blocks,512>>>(out,1.0f,cyc); cudaDeviceSynchronize();
  long long h[512]; cudaMemcpy(h,cyc,sizeof(long long)*blocks,cudaMemcpyDeviceToHost);
  double avg=0; for(int i=0;i<blocks;++i) avg+=h[i]; avg/=blocks;
  const double winstr_per_sm = 16.0 /*warps*/ * ITER * REP * NCH * instr_per_chain;
  printf("%-28s cycles %9.0f  warp-instr/clk/SM %.3f   (per SMSP %.3f)\n", name, avg, winstr_per_sm/avg, winstr_per_sm/avg/4);
  cudaFree(out); cudaFree(cyc);
}

int main(){
  run<0>("FFMA",1); run<1>("FFMA2",1); run<15>("FMUL2",1); run<8>("FADD2",1); run<9>("FADD",1);
  run<2>("PRMT",1); run<3>("LOP3",1); run<4>("FMNMX",1); run<10>("IADD3",1);
  run<5>("FFMA+PRMT",2); run<6>("FFMA2+PRMT",2); run<7>("FFMA2+FFMA",2);
  run<12>("FFMA2+PRMT+LOP/SHF",3); run<21>("FFMA2+PRMT+LOP3 indep",3);
  run<16>("I2F.S8 chain",1); run<17>("I2F.S8+FADD",2); run<18>("I2F.S8+FADD+PRMT",3); run<19>("I2F.S8+3FFMA",4); run<20>("4FFMA+PRMT",5); run<13>("2FFMA+PRMT",3); run<14>("FFMA+LOP3+PRMT",3);
  cudaError_t e=cudaGetLastError(); if(e!=cudaSuccess){printf("err %s\n",cudaGetErrorString(e));return 1;}
  return 0;
}
