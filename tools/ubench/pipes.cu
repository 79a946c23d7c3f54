// Issue rate of the integer instructions the ray march is made of, on this GPU (B200, sm_100a): warp instructions per
// clock per SM for IADD3, LOP3, SHF, ISETP+SEL, IMAD, IMAD.HI, IMAD.WIDE -- the facts behind "the march is bound by ...".
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) k(unsigned *out, unsigned a0, unsigned b0)
{
  unsigned x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = a0 + threadIdx.x * (i + 1);
  unsigned b = b0 | 1u, c = b0 * 3u + 7u;
  for (int it = 0; it < ITERS; it++)
  {
#pragma unroll
    for (int i = 0; i < ILP; i++)
    {
      if (OP == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                                  // IADD3
      if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));              // LOP3
      if (OP == 2) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));              // SHF
      if (OP == 3) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(x[i]) : "r"(b), "r"(c));   // ISETP + SEL
      if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));                  // IMAD
      if (OP == 5) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                              // IMAD.HI.U32
      if (OP == 6) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                              // IMAD.HI
      if (OP == 7) asm volatile("{.reg .u64 w; mul.wide.u32 w, %0, %1; cvt.u32.u64 %0, w;}" : "+r"(x[i]) : "r"(b));   // IMAD.WIDE (low word used)
      if (OP == 8) asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; @p sub.u32 %0, %0, %1;}" : "+r"(x[i]) : "r"(b));   // ISETP + predicated IADD
      if (OP == 9) asm volatile("popc.b32 %0, %0;" : "+r"(x[i]));                                              // POPC
      if (OP == 10) asm volatile("{.reg .f32 f; cvt.rn.f32.u32 f, %0; mov.b32 %0, f;}" : "+r"(x[i]));          // I2F
      if (OP == 11) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[i]) : "r"(b));                           // ptxas turns it into IADD3
      if (OP == 12) asm volatile("{.reg .u32 t; shr.u32 t, %0, 31; add.u32 %0, t, %1;}" : "+r"(x[i]) : "r"(b));   // LEA.HI (shift + add)
      if (OP == 13) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(x[i]) : "r"(b));                        // PRMT
      if (OP == 14) asm volatile("{.reg .pred p; setp.ne.u32 p, %0, %1; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(x[i]) : "r"(b));   // ISETP + VOTE
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= x[i];
  if (s == 0x12345678u) out[0] = s;
}

template <int OP>
static void run(const char *name, int ops_per_iter, int sms, double mhz)
{
  unsigned *d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int ctas = sms * 8;      // 64 warps per SM resident
  k<OP><<<ctas, 256>>>(d, 1u, 2u);
  cudaEventRecord(e0);
  k<OP><<<ctas, 256>>>(d, 1u, 2u);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)ctas * 8 * ITERS * ILP * ops_per_iter;
  const double clocks = ms * 1e-3 * mhz * 1e6;
  printf("%-28s %7.3f ms  %6.2f warp-instr / clk / SM  (%.2f per scheduler)\n", name, ms, warp_instr / clocks / sms, warp_instr / clocks / sms / 4);
  cudaFree(d);
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  printf("%s, %d SMs, %.0f MHz (nominal; rates assume it)\n", p.name, p.multiProcessorCount, mhz);
  const int s = p.multiProcessorCount;
  run<0>("IADD3", 1, s, mhz); run<1>("LOP3", 1, s, mhz); run<2>("SHF", 1, s, mhz); run<3>("ISETP + SEL", 2, s, mhz);
  run<8>("ISETP + @p IADD", 2, s, mhz); run<4>("IMAD", 1, s, mhz); run<5>("IMAD.HI.U32", 1, s, mhz); run<6>("IMAD.HI (signed)", 1, s, mhz);
  run<7>("IMAD.WIDE", 1, s, mhz); run<9>("POPC", 1, s, mhz); run<10>("I2F", 1, s, mhz);
  run<11>("mad x, 1, b (ptxas: IADD3)", 1, s, mhz); run<12>("LEA.HI (x >> 31) + b", 1, s, mhz); run<13>("PRMT", 1, s, mhz);
  run<14>("ISETP + VOTE", 2, s, mhz);
  return 0;
}
