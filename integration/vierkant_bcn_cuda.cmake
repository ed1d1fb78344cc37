# vierkant_bcn_cuda.cmake -- lives beside cmake_modules/build_shaders.cmake in vierkant.
#
#   option(VIERKANT_BCN_CUDA "BC7/BC5 texture compression on NVIDIA B200 (sm_100a) instead of the CPU" OFF)
#   include(vierkant_bcn_cuda)            # after the vierkant target exists
#
# With the option ON, src/texture_block_compression.cpp is taken out of the vierkant target and replaced by
# texture_block_compression_cuda.cpp (same vierkant::bcn::compress symbol), and libvierkant_bcn_cuda is built from the
# CUDA sources with the numerics flags bit-exactness depends on (no FMA contraction, IEEE division and square root).
option(VIERKANT_BCN_CUDA "BC7/BC5 texture compression on NVIDIA B200 (sm_100a) instead of the CPU" OFF)

if(VIERKANT_BCN_CUDA)
    enable_language(CUDA)
    set(VIERKANT_BCN_CUDA_DIR "${CMAKE_CURRENT_LIST_DIR}/.." CACHE PATH "checkout of the vierkant-bcn-b200 repository")

    add_library(vierkant_bcn_cuda SHARED
            ${VIERKANT_BCN_CUDA_DIR}/vierkant_b200/csrc/bcn_cuda.cu
            ${VIERKANT_BCN_CUDA_DIR}/vierkant_b200/csrc/bc7_tables.cpp)
    target_include_directories(vierkant_bcn_cuda PUBLIC ${VIERKANT_BCN_CUDA_DIR}/include)
    set_target_properties(vierkant_bcn_cuda PROPERTIES CUDA_ARCHITECTURES "100a" CUDA_STANDARD 17 POSITION_INDEPENDENT_CODE ON)
    target_compile_options(vierkant_bcn_cuda PRIVATE
            $<$<COMPILE_LANGUAGE:CUDA>:-fmad=false -prec-div=true -prec-sqrt=true -ftz=false -lineinfo -Xcompiler=-ffp-contract=off>
            $<$<COMPILE_LANGUAGE:CXX>:-ffp-contract=off>)

    # swap the translation unit that defines vierkant::bcn::compress
    get_target_property(_vkt_sources vierkant SOURCES)
    list(FILTER _vkt_sources EXCLUDE REGEX "texture_block_compression\\.cpp$")
    set_target_properties(vierkant PROPERTIES SOURCES "${_vkt_sources}")
    target_sources(vierkant PRIVATE ${VIERKANT_BCN_CUDA_DIR}/integration/texture_block_compression_cuda.cpp)
    # texture_block_compression_batch.hpp: the optional several-textures-per-call overload (for model::compress_textures)
    target_include_directories(vierkant PUBLIC ${VIERKANT_BCN_CUDA_DIR}/integration)
    target_link_libraries(vierkant PUBLIC vierkant_bcn_cuda)
    target_compile_definitions(vierkant PUBLIC VIERKANT_BCN_CUDA=1)
endif()
