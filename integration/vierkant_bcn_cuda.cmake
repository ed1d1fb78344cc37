# vierkant_bcn_cuda.cmake -- lives beside cmake_modules/build_shaders.cmake in vierkant (integration/vierkant.patch shows the
# three places it is used from).
#
#   option VIERKANT_BCN_CUDA          BC7/BC5 texture compression on NVIDIA B200 (sm_100a) instead of the CPU
#   vierkant_bcn_cuda_sources(<var>)   before add_library(vierkant ...): in the source list <var>, replace
#                                      src/texture_block_compression.cpp by texture_block_compression_cuda.cpp (same
#                                      vierkant::bcn::compress symbol) -- the target is never edited behind its owner's back
#   vierkant_bcn_cuda_link(<target>)   after add_library: link libvierkant_bcn_cuda, add the batch-overload header
#
# libvierkant_bcn_cuda itself is built here from the CUDA sources with the numerics flags bit-exactness depends on (no FMA
# contraction, IEEE division and square root, no flush-to-zero).
option(VIERKANT_BCN_CUDA "BC7/BC5 texture compression on NVIDIA B200 (sm_100a) instead of the CPU" OFF)

if(VIERKANT_BCN_CUDA)
    set(VIERKANT_BCN_CUDA_DIR "${CMAKE_CURRENT_LIST_DIR}/.." CACHE PATH "checkout of the vierkant-bcn-b200 repository")
    if(NOT DEFINED CMAKE_CUDA_ARCHITECTURES)
        set(CMAKE_CUDA_ARCHITECTURES "100a")
    endif()
    enable_language(CUDA)

    add_library(vierkant_bcn_cuda SHARED
            ${VIERKANT_BCN_CUDA_DIR}/vierkant_b200/csrc/bcn_cuda.cu
            ${VIERKANT_BCN_CUDA_DIR}/vierkant_b200/csrc/bc7_tables.cpp)
    target_include_directories(vierkant_bcn_cuda PUBLIC ${VIERKANT_BCN_CUDA_DIR}/include)
    set_target_properties(vierkant_bcn_cuda PROPERTIES CUDA_ARCHITECTURES "100a" CUDA_STANDARD 17 POSITION_INDEPENDENT_CODE ON)
    target_compile_options(vierkant_bcn_cuda PRIVATE
            $<$<COMPILE_LANGUAGE:CUDA>:-fmad=false -prec-div=true -prec-sqrt=true -ftz=false -lineinfo -Xcompiler=-ffp-contract=off>
            $<$<COMPILE_LANGUAGE:CXX>:-ffp-contract=off>)
endif()

function(vierkant_bcn_cuda_sources SOURCES_VAR)
    if(NOT VIERKANT_BCN_CUDA)
        return()
    endif()
    set(_sources ${${SOURCES_VAR}})
    list(FILTER _sources EXCLUDE REGEX "(^|/)texture_block_compression\\.cpp$")
    list(APPEND _sources ${VIERKANT_BCN_CUDA_DIR}/integration/texture_block_compression_cuda.cpp)
    set(${SOURCES_VAR} ${_sources} PARENT_SCOPE)
endfunction()

function(vierkant_bcn_cuda_link TARGET)
    if(NOT VIERKANT_BCN_CUDA)
        return()
    endif()
    # texture_block_compression_batch.hpp: the several-textures-per-call overload model::compress_textures uses
    target_include_directories(${TARGET} PUBLIC ${VIERKANT_BCN_CUDA_DIR}/integration)
    target_link_libraries(${TARGET} PUBLIC vierkant_bcn_cuda)
    target_compile_definitions(${TARGET} PUBLIC VIERKANT_BCN_CUDA=1)
endfunction()
