// Throughput of the integer instructions the reconstruction kernels lean on (sm_100a), and the overflow
// semantics of VIADDMNMX.S16x2.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 int_pipes.cu -o int_pipes
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#define ITER 4096
template <int OP> __global__ void __launch_bounds__(256) k(int *out, int a0, int b0)
{
    int a[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = a0 + threadIdx.x + i; c[i] = b0 + i; }
    int b = b0 | 0x01020304;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) c[i] = a[i] * b + c[i];                                  // IMAD
            if (OP == 1) c[i] = __dp2a_lo(a[i], b, c[i]);                        // IDP.2A
            if (OP == 2) c[i] = __dp4a(a[i], b, c[i]);                           // IDP.4A
            if (OP == 3) c[i] = __byte_perm(c[i], a[i], 0x5410) ^ b;             // PRMT + LOP3
            if (OP == 4) c[i] = __viaddmin_s16x2_relu(c[i], a[i], b);            // VIADDMNMX.S16x2.RELU
            if (OP == 5) c[i] = (c[i] >> 3) + a[i];                              // SHF + IADD
            if (OP == 6) asm volatile("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(c[i]) : "r"(c[i]), "r"(a[i]));   // I2IP
            if (OP == 7) c[i] = __float_as_int(__int_as_float(c[i]) * 1.0001f + 3.0f);   // FFMA imm
            if (OP == 8) c[i] = max(min(c[i] + a[i], b), 0);                     // IADD + VIMNMX x2 (or VIADDMNMX 32)
            if (OP == 9) { c[i] = a[i] * b + c[i]; a[i] = __byte_perm(a[i], c[i], 0x1032); }   // IMAD + PRMT dual-issue check
            if (OP == 10) { c[i] = __dp2a_lo(a[i], b, c[i]); a[i] = (a[i] >> 1) ^ c[i]; }       // IDP + SHF/LOP
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i] + a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void sem(int *o)
{
    // 0x7fff + 1 per halfword: wraps to 0x8000 (then relu -> 0) or saturates (min with c)?
    o[0] = __viaddmin_s16x2_relu(0x7fff7fff, 0x00010001, 0x03ff03ff);
    o[1] = __viaddmin_s16x2_relu(0x7f007f00, 0x03ff03ff, 0x03ff03ff);
    o[2] = __viaddmin_s16x2_relu(0x80008000, 0xffffffff, 0x03ff03ff);   // -32768 + -1
}

int main()
{
    int *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const char *names[] = {"IMAD", "IDP.2A", "IDP.4A", "PRMT+LOP3", "VIADDMNMX.S16x2.RELU", "SHF+IADD", "I2IP.SAT", "FFMA", "IADD+2xVIMNMX", "IMAD+PRMT", "IDP.2A+SHF+LOP"};
    const int nops[] = {1, 1, 1, 2, 1, 2, 1, 1, 3, 2, 3};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock attr %d kHz\n", p.name, p.multiProcessorCount, clk);
    for (int op = 0; op <= 10; op++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            switch (op) {
            case 0: k<0><<<148 * 8, 256>>>(d, 1, 2); break;  case 1: k<1><<<148 * 8, 256>>>(d, 1, 2); break;
            case 2: k<2><<<148 * 8, 256>>>(d, 1, 2); break;  case 3: k<3><<<148 * 8, 256>>>(d, 1, 2); break;
            case 4: k<4><<<148 * 8, 256>>>(d, 1, 2); break;  case 5: k<5><<<148 * 8, 256>>>(d, 1, 2); break;
            case 6: k<6><<<148 * 8, 256>>>(d, 1, 2); break;  case 7: k<7><<<148 * 8, 256>>>(d, 1, 2); break;
            case 8: k<8><<<148 * 8, 256>>>(d, 1, 2); break;  case 9: k<9><<<148 * 8, 256>>>(d, 1, 2); break;
            case 10: k<10><<<148 * 8, 256>>>(d, 1, 2); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 1) {
                double ops = 148.0 * 8 * 256 * (double)ITER * 8;
                printf("%-24s %8.3f ms  %7.2f Gop-lines/s  = %6.1f source-ops/clk/SM at 1.9 GHz (x%d SASS)\n", names[op], ms, ops / ms / 1e6,
                       ops / (ms * 1e-3) / 148 / 1.9e9, nops[op]);
            }
        }
    }
    int h[3]; sem<<<1, 1>>>(d); cudaMemcpy(h, d, 12, cudaMemcpyDeviceToHost);
    printf("viaddmin_s16x2_relu(0x7fff+1, c=0x3ff) = %08x ; (0x7f00+0x3ff) = %08x ; (-32768 + -1) = %08x\n", h[0], h[1], h[2]);
    return 0;
}
