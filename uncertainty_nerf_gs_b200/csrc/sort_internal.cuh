// Host-side entry points of radix_sort.cu shared with select_cuts.cu (same library, not exported).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace ub {

size_t segmented_sort_workspace(int num_segments, long long total, long long max_len, bool with_vals);
size_t cut_prefix_workspace(int num_segments, long long max_len, int num_values, int num_cuts);

// ub_segmented_sort with an optional DEVICE table of used lengths: segment s occupies the slot
// [off[s], off[s+1]) but only its first seg_lens[s] keys are sorted (grids are sized for the slots).
int segmented_sort_impl(const float* keys, int32_t num_segments, const int64_t* seg_offsets,
                        const int64_t* seg_lens, int64_t total, int64_t max_segment_len, float* out_sorted_keys,
                        int32_t* out_perm, void* workspace, size_t workspace_bytes, void* stream_v);

// ub_cut_prefix_sums with optional per-value length / cut tables (value stride in elements, 0 = shared) and
// an optional addend of the output's shape.
int cut_prefix_impl(const float* const* values_host, const int32_t* const* perms_host, int32_t num_values,
                    int32_t num_segments, const int64_t* seg_offsets, const int64_t* seg_lens,
                    int64_t lens_value_stride, int64_t max_segment_len, const int64_t* cuts,
                    int64_t cuts_value_stride, int32_t num_cuts, const double* add, double* out_sums,
                    void* workspace, size_t workspace_bytes, void* stream_v);

}  // namespace ub
