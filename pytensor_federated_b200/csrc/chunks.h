// Chunk table of the dynamically scheduled tensor-core GLM kernels (host side; glm_tc.cu, glm_fp8.cu).
#pragma once
#include <vector>

#include "models.h"

// Consecutive tiles of one segment, a multiple of `multiple` of them (the epilogue groups of a kernel rotate over
// the tiles, so a chunk must hold a whole number of rotations: 2 for the bf16 kernel, 3 for the fp8 kernel; a
// segment whose tile count is not such a multiple ends in empty tiles), `max_chunk` tiles while plenty of work
// is left and shrinking towards `min_chunk` as the remaining work approaches two chunks per SM (guided
// self-scheduling), so the last chunks that are handed out are the small ones.
inline std::vector<GlmChunk> build_chunks(const GlmSegment* segs, int n_segments, int sm_count, int tile_rows, int multiple,
                                          int max_chunk, int min_chunk) {
    std::vector<GlmChunk> out;
    long long remaining = 0;
    for (int s = 0; s < n_segments; ++s) remaining += (segs[s].n_rows + tile_rows - 1) / tile_rows;
    if (remaining <= 8ll * sm_count) min_chunk = multiple;   // small problems: spread over the SMs first
    for (int s = 0; s < n_segments; ++s) {
        const long long ts = (segs[s].n_rows + tile_rows - 1) / tile_rows;
        long long t = 0;
        while (t < ts) {
            long long want = remaining / (2ll * sm_count);
            want -= want % multiple;
            if (want < min_chunk) want = min_chunk;
            if (want > max_chunk) want = max_chunk;
            const long long real = want < ts - t ? want : ts - t;
            const long long n = (real + multiple - 1) / multiple * multiple;
            out.push_back(GlmChunk{s, (int)t, (int)n, 0});
            t += real;
            remaining -= real;
        }
    }
    return out;
}
