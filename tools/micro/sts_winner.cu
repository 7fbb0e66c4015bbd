// Which lane's value survives when several lanes of one STS write the same shared-memory address? (sm_100a probe)
#include <cstdio>
#include <cstdint>
__global__ void k(int* out, int pattern)
{
    __shared__ volatile uint16_t s16[64];
    __shared__ volatile uint32_t s32[64];
    const int lane = threadIdx.x;
    int idx;
    switch (pattern) {
    case 0: idx = 0; break;                 // all lanes, one address
    case 1: idx = lane & 1; break;          // two sets, interleaved
    case 2: idx = lane >> 3; break;         // four sets of 8 neighbours
    case 3: idx = (lane * 7) & 3; break;    // scattered
    default: idx = lane % 5; break;
    }
    s16[idx] = (uint16_t)lane;
    s32[idx] = (uint32_t)lane;
    __syncwarp();
    out[lane] = s16[idx];
    out[32 + lane] = (int)s32[idx];
}
int main()
{
    int* d; cudaMalloc(&d, 64 * sizeof(int));
    for (int p = 0; p < 5; p++) {
        k<<<1, 32>>>(d, p);
        int h[64]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("pattern %d u16:", p); for (int i = 0; i < 32; i++) printf(" %d", h[i]); printf("\n");
        printf("pattern %d u32:", p); for (int i = 0; i < 32; i++) printf(" %d", h[32 + i]); printf("\n");
    }
    return 0;
}
