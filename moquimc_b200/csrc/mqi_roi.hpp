// mqi_roi.hpp -- run-length regions of interest, host side.
//
// mask_reader::mask_to_roi (moqui/base/mqi_file_handler.hpp:176-217) turns the summed 0/1 mask volumes
// (read_mask_files adds the masks of all files, :107-113) into a CONTOUR roi_t: parallel arrays start /
// stride / accumulated stride of the runs of consecutive voxels (moqui/base/mqi_roi.hpp:20-45).  The
// transport kernel only asks "is voxel cnb inside a run" (roi_t::idx -> idx_contour, :127-137), the
// writers ask for the compressed index (get_mask_idx -> get_contour_idx, :88-100) and the roi size
// (get_mask_size = acc_stride[length-1], :75-86).
//
// Reference behaviour kept: a run opens at a voxel whose summed mask is exactly 1 and closes at the
// next voxel whose sum is 0, so voxels where two masks overlap (sum >= 2) neither open nor close a run.
// Deviation (documented): a run still open at the end of the volume has no stride in the reference
// (the vectors end up with different lengths and uninitialised memory is read); here it is closed at
// the end of the volume.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace mqib
{
struct RoiRuns {
    std::vector<uint32_t> start, stride, acc_stride;
    uint64_t              original_length = 0;

    uint32_t size() const { return acc_stride.empty() ? 0u : acc_stride.back(); }   // get_mask_size()

    // get_contour_idx: compressed index of voxel v, or -1 outside every run (binary search like lower_bound_cpp)
    int64_t
    compressed_index(uint32_t v) const {
        size_t lo = 0, hi = start.size();
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            if (start[mid] <= v) lo = mid + 1; else hi = mid;
        }
        if (lo == 0) return -1;
        const size_t   c = lo - 1;
        const uint32_t d = v - start[c];
        if (d >= stride[c]) return -1;
        return (int64_t) d + (c > 0 ? acc_stride[c - 1] : 0u);
    }

    // one bit per voxel, 32 voxels per word (the kernel-side form)
    std::vector<uint32_t>
    bitmask() const {
        std::vector<uint32_t> bits((original_length + 31) / 32, 0u);
        for (size_t r = 0; r < start.size(); ++r)
            for (uint64_t v = start[r]; v < (uint64_t) start[r] + stride[r]; ++v) bits[v >> 5] |= 1u << (v & 31);
        return bits;
    }
};

inline RoiRuns
mask_to_roi(const uint8_t* mask_total, uint64_t n) {
    RoiRuns  r;
    bool     open = false;
    uint64_t s0   = 0;
    r.original_length = n;
    auto close_run = [&](uint64_t end) {
        r.start.push_back((uint32_t) s0);
        r.stride.push_back((uint32_t) (end - s0));
        r.acc_stride.push_back((r.acc_stride.empty() ? 0u : r.acc_stride.back()) + (uint32_t) (end - s0));
    };
    for (uint64_t i = 0; i < n; ++i) {
        if (mask_total[i] == 1 && !open) {
            open = true;
            s0   = i;
        }
        if (mask_total[i] == 0 && open) {
            open = false;
            close_run(i);
        }
    }
    if (open) close_run(n);
    return r;
}
}   // namespace mqib
