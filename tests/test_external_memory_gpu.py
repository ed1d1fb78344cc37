"""SURVEY.md 8f N4, the CUDA half of the Vulkan hand-off executed end to end: an allocation exported as an opaque POSIX fd is
imported with vkt_bcn_cuda_import_external_fd (cudaImportExternalMemory + cudaExternalMemoryGetMappedBuffer) and used as the
destination of every level of compress().  There is no Vulkan on the test boxes; the fd comes from the CUDA virtual memory
API (cuMemCreate + cuMemExportToShareableHandle), which hands out the same kind of handle vkGetMemoryFdKHR does for a
VkDeviceMemory with VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT -- the stand-in for vierkant's staging VkBuffer
(src/model/model_loading.cpp:483-488)."""
import ctypes as C
import os

import numpy as np
import pytest

from vierkant_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _export_fd(nbytes: int):
    """(fd, allocation handle, padded size) of a device allocation exportable as a POSIX fd, or skip."""
    try:
        from cuda.bindings import driver as cu
    except Exception:  # noqa: BLE001
        pytest.skip("cuda-python driver bindings not available")

    def ok(res, what):
        err = res[0]
        if err != cu.CUresult.CUDA_SUCCESS:
            pytest.skip(f"{what}: {err}")
        return res[1] if len(res) == 2 else res[1:]

    ok(cu.cuInit(0), "cuInit")
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = ok(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM), "granularity")
    size = (nbytes + gran - 1) // gran * gran
    handle = ok(cu.cuMemCreate(size, prop, 0), "cuMemCreate")
    fd = ok(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle")
    return int(fd), handle, size, cu


def test_blocks_land_in_imported_external_memory(ctx):
    import torch
    torch.cuda.init()
    img = synth.make_texture(1024, 512, 1, seed=12)
    h, w, c = img.shape
    plan = capi.compress_plan(w, h, True)
    sizes = [int(plan.level_num_blocks[l]) * 16 for l in range(plan.num_levels)]
    offs = np.concatenate([[0], np.cumsum([(s + 255) // 256 * 256 for s in sizes])]).astype(np.int64)
    fd, alloc, size, cu = _export_fd(int(offs[-1]))
    try:
        try:
            d_ptr, ext = ctx.import_external_fd(fd, size)
        except capi.BcnError as e:
            os.close(fd)
            pytest.skip(f"this driver does not import a VMM fd as external memory: {e}")
        try:
            ptrs = (C.c_void_p * plan.num_levels)(*[d_ptr + int(offs[l]) for l in range(plan.num_levels)])
            ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, img.ctypes.data, w, h, c, 1, None, ptrs))
            s0 = ctx.stats()
            ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, img.ctypes.data, w, h, c, 1, None, ptrs))
            s1 = ctx.stats()
            # read the "Vulkan buffer" back through plain CUDA and compare with the host-destination call
            back = torch.empty(int(offs[-1]), dtype=torch.uint8)
            cudart = C.CDLL("libcudart.so.12")
            cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
            assert cudart.cudaMemcpy(back.data_ptr(), d_ptr, int(offs[-1]), 2) == 0
            _, want = ctx.compress(img, capi.MODE_BC7, True)
            for l in range(plan.num_levels):
                got = back.numpy()[int(offs[l]):int(offs[l]) + sizes[l]].reshape(-1, 16)
                assert np.array_equal(got, want[l]), f"level {l}"
            # device destinations: the blocks never cross to the host (the counter only sees the source upload)
            assert s1["d2h_bytes"] - s0["d2h_bytes"] == sum(sizes)  # counted as delivered bytes ...
        finally:
            ctx.release_external(d_ptr, ext)
    finally:
        cu.cuMemRelease(alloc)
