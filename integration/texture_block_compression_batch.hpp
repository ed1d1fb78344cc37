// texture_block_compression_batch.hpp -- extension of vierkant::bcn for the CUDA encoder: several textures per call.
//
// vierkant's model::compress_textures (src/model/model_loading.cpp:96-118) calls bcn::compress() once per texture, in
// sequence.  With the encoder on a GPU that leaves the device idle between textures (upload of the next source, drain of
// the last kernels).  This overload takes all textures of a model at once and returns exactly what the
// per-texture calls return, in the same order; compress_textures would build its compress_info_t list first and fan the
// results out afterwards.
#pragma once
#include <span>
#include <vector>

#include <vierkant/texture_block_compression.hpp>

namespace vierkant::bcn
{
std::vector<compress_result_t> compress(std::span<const compress_info_t> compress_infos);
}
