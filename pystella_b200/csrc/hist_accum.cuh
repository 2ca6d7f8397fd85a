// Warp-aggregated histogram accumulation, shared by the hand-written spectrum binning
// kernel (diag.cu) and the run-time generated Histogrammer kernels (codegen.py inlines
// this file into the NVRTC source, so: no host headers, no dependencies).
//
// pystella's Histogrammer (pystella/histogram.py:60-170) accumulates
//     temp[j, floor(bin_expr)] += weight_expr          (work-group local, atomic)
//     hist[j, b] += temp[j, b]                         (global, atomic)
// The same two levels here, with one more in front: the 32 lanes of a warp hold
// consecutive z sites, whose bins often coincide (always, for the |k| shells of a power
// spectrum), so equal neighbours are first summed with shuffles and only the head of each
// run issues the shared-memory atomic.
#pragma once

namespace psb {

// Every lane of the warp must call this (full-mask shuffles); lanes without a site pass
// bin = -1.  `bins` points to num_bins doubles (shared or global memory).
__device__ __forceinline__ void hist_accumulate(double* bins, int num_bins, int bin,
                                                double w, int lane) {
  const unsigned full = 0xffffffffu;
  if (bin < 0 || bin >= num_bins) {
    bin = -1;
    w = 0.0;
  }
  const int prev = __shfl_up_sync(full, bin, 1);
  const bool head = (lane == 0) || (prev != bin);
  const unsigned heads = __ballot_sync(full, head);
  // first head above this lane = end of the run this lane belongs to
  const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
  const int seg_end = above ? lane + __ffs(above) : 32;
  // suffix sums clipped to the run (head lane ends up with the run's total)
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double w2 = __shfl_down_sync(full, w, d);
    if (lane + d < seg_end) w += w2;
  }
  if (head && bin >= 0) atomicAdd(&bins[bin], w);
}

// Flush `copies` block-private copies of `n` bins into the global histogram.
__device__ __forceinline__ void hist_flush(const double* temp, int copies, int n, double* hist,
                                           int tid, int nthreads) {
  for (int b = tid; b < n; b += nthreads) {
    double s = 0.0;
    for (int c = 0; c < copies; ++c) s += temp[c * n + b];
    if (s != 0.0) atomicAdd(&hist[b], s);
  }
}

}  // namespace psb
