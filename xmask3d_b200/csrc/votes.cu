// Cross-view vote accumulation — the step right after the path at inference:
//   scene_pred[mask_2d, logits_pred] += 1 ; counter[mask_2d] += 1      (reference run/infer.py:642-647)
//   _, scene_pred = torch.max(scene_pred, dim=1)                       (run/infer.py:658)
// batched over all views of a batch of scenes: every visible (view, point) pair adds one vote for its
// predicted class to its scene point.  Integer atomics: the result does not depend on the order.
#include "common.cuh"

namespace xm3d {

__global__ void __launch_bounds__(256)
vote_kernel(const int32_t *__restrict__ vis_idx, const int64_t *__restrict__ seg_off, int n_seg, int64_t cap,
            const int64_t *__restrict__ view_pt_off, const int32_t *__restrict__ cls, int n_classes,
            int32_t *__restrict__ votes, int32_t *__restrict__ counter) {
    int64_t total = seg_off[n_seg];
    if (total > cap) total = 0;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total) return;
    const int s = seg_of(seg_off, n_seg, j);
    const int64_t p = view_pt_off[s] + vis_idx[j];
    const int c = cls[j];
    if (c >= 0 && c < n_classes) atomicAdd(&votes[p * n_classes + c], 1);
    atomicAdd(&counter[p], 1);
}

// first maximum of every row (torch.max semantics on ties); -1 for points no view has seen
__global__ void __launch_bounds__(256)
vote_argmax_kernel(const int32_t *__restrict__ votes, const int32_t *__restrict__ counter, int64_t n_pts,
                   int n_classes, int32_t *__restrict__ pred) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const int32_t *row = votes + p * n_classes;
    int best = 0, best_v = row[0];
    for (int c = 1; c < n_classes; ++c) {
        const int v = row[c];
        if (v > best_v) { best_v = v; best = c; }
    }
    pred[p] = counter[p] > 0 ? best : -1;
}

}  // namespace xm3d

using namespace xm3d;

extern "C" int xm3d_vote_batch(const int32_t *vis_idx, const int64_t *seg_off, int32_t n_seg, int64_t cap,
                               const int64_t *view_pt_off, const int32_t *cls, int32_t n_classes, int32_t *votes,
                               int32_t *counter, xm3d_stream_t stream) {
    XM3D_REQUIRE(n_seg > 0 && cap >= 0 && n_classes > 0, "bad sizes");
    XM3D_REQUIRE(vis_idx && seg_off && view_pt_off && cls && votes && counter, "null pointer");
    if (cap == 0) return XM3D_OK;
    vote_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        vis_idx, seg_off, n_seg, cap, view_pt_off, cls, n_classes, votes, counter);
    count_launches(1);
    return check_launch("xm3d_vote_batch");
}

extern "C" int xm3d_vote_argmax(const int32_t *votes, const int32_t *counter, int64_t n_pts, int32_t n_classes,
                                int32_t *pred, xm3d_stream_t stream) {
    XM3D_REQUIRE(n_pts >= 0 && n_classes > 0, "bad sizes");
    if (n_pts == 0) return XM3D_OK;
    XM3D_REQUIRE(votes && counter && pred, "null pointer");
    vote_argmax_kernel<<<(unsigned)((n_pts + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        votes, counter, n_pts, n_classes, pred);
    count_launches(1);
    return check_launch("xm3d_vote_argmax");
}
