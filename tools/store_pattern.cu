// Micro-benchmark: achieved HBM write (and read-modify-write) bandwidth for tile-shaped access to a row-major
// [3072, 162816] matrix (the out_layer.fc1 weight / gradient), as a function of the contiguous bytes per row that one
// tile covers and of the tile visiting order.  Explains why a 128x128-tile GEMM epilogue writes at ~2 TB/s while a
// linear AdamW pass reaches ~7 TB/s, and what tile width the fused wgrad+AdamW kernel needs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/store_pattern tools/store_pattern.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

// One CTA = 16 warps.  Tile = 128 rows x tile_w bytes.  Warp w handles rows (w%4)*32..+32 and the (w/4)-th quarter of
// the tile width (mirrors the GEMM epilogue's TMEM quadrant / column-part split).  seg = contiguous bytes per row per
// warp store instruction (16 B per lane).
template <int MODE, int U>  // MODE 0 = write only, 1 = read + write (same address), 2 = read only; U = independent
                             // 16-byte accesses a lane keeps in flight (U = 1 reproduces the round-1 table, whose
                             // read / rmw rows were limited by loads in flight, not by the access pattern)
__global__ void __launch_bounds__(512, 1)
tile_kernel(char* base, long long pitch, int m_tiles, int n_tiles, int tile_w, int seg, int raster, int esz_shift,
            unsigned long long* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, part = warp >> 2;
  const int total = m_tiles * n_tiles;
  int w_begin, w_end, w_step;
  if (raster == 0) { w_begin = blockIdx.x; w_end = total; w_step = gridDim.x; }
  else {
    const int per = (total + gridDim.x - 1) / gridDim.x;
    w_begin = blockIdx.x * per; w_end = min(total, w_begin + per); w_step = 1;
  }
  const int lanes_per_row = seg / 16, rows_per_inst = 32 / lanes_per_row;
  const int part_w = tile_w / 4;
  const int insts_per_c = 32 / rows_per_inst;            // store instructions per seg-wide column step
  const int n_inst = (part_w / seg) * insts_per_c;       // per tile and warp
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int w = w_begin; w < w_end; w += w_step) {
    const int mt = raster ? w / n_tiles : w % m_tiles, nt = raster ? w % n_tiles : w / m_tiles;
    char* tile = base + (long long)(mt * 128 + quad * 32) * pitch + (long long)nt * tile_w + part * part_w;
    for (int i0 = 0; i0 < n_inst; i0 += U) {
      char* p[U];
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = min(i0 + u, n_inst - 1);
        const int c = (i / insts_per_c) * seg, r = (i % insts_per_c) * rows_per_inst;
        p[u] = tile + (long long)(r + lane / lanes_per_row) * pitch + c + (lane % lanes_per_row) * 16;
      }
      if (MODE != 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = *reinterpret_cast<const uint4*>(p[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i0 + u >= n_inst) break;
        if (MODE == 0) {
          *reinterpret_cast<uint4*>(p[u]) = make_uint4(w, i0, u, lane);
        } else if (MODE == 1) {
          v[u].x += 1;
          *reinterpret_cast<uint4*>(p[u]) = v[u];
        } else {
          acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w;
        }
      }
    }
  }
  if (MODE == 2 && (acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) *sink = 1;
}

template <int U>
static void launch(int mode, char* buf, long long pitch, int m_tiles, int n_tiles, int tw, int seg, int raster,
                   unsigned long long* sink) {
  if (mode == 0) tile_kernel<0, U><<<148, 512>>>(buf, pitch, m_tiles, n_tiles, tw, seg, raster, 0, sink);
  else if (mode == 1) tile_kernel<1, U><<<148, 512>>>(buf, pitch, m_tiles, n_tiles, tw, seg, raster, 0, sink);
  else tile_kernel<2, U><<<148, 512>>>(buf, pitch, m_tiles, n_tiles, tw, seg, raster, 0, sink);
}

int main() {
  const long long rows = 3072, cols_bytes = 162816LL * 4;   // fp32 matrix, 2.0 GB
  char* buf;
  unsigned long long* sink;
  cudaMalloc(&buf, rows * cols_bytes);
  cudaMalloc(&sink, 8);
  cudaMemset(buf, 0, rows * cols_bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%-6s %-7s %-6s %-6s %-4s %10s %10s\n", "mode", "tile_w", "seg", "raster", "U", "us", "GB/s");
  const int tws[] = {256, 512, 1024, 2048, 4096, 8192};
  for (int mode = 0; mode < 3; ++mode)
    for (int ti = 0; ti < 6; ++ti)
      for (int raster = 0; raster < 2; ++raster)
        for (int seg = 64; seg <= 512; seg *= 2)
         for (int U = 1; U <= 16; U *= 4) {
          const int tw = tws[ti];
          if (mode == 0 && U != 1) continue;
          if (seg > tw / 4) continue;
          if (seg != 128 && !(tw == 512 || tw == 2048)) continue;   // sweep seg only at two widths
          const int m_tiles = rows / 128, n_tiles = (int)(cols_bytes / tw);
          float best = 1e30f;
          for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            if (U == 1) launch<1>(mode, buf, cols_bytes, m_tiles, n_tiles, tw, seg, raster, sink);
            else if (U == 4) launch<4>(mode, buf, cols_bytes, m_tiles, n_tiles, tw, seg, raster, sink);
            else launch<16>(mode, buf, cols_bytes, m_tiles, n_tiles, tw, seg, raster, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
          }
          const double bytes = (double)m_tiles * 128 * n_tiles * tw * (mode == 1 ? 2 : 1);
          printf("%-6s %-7d %-6d %-6d %-4d %10.1f %10.1f\n", mode == 0 ? "write" : (mode == 1 ? "rmw" : "read"), tw,
                 seg, raster, U, best * 1e3, bytes / best / 1e6);
          fflush(stdout);
        }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
