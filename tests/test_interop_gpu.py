"""The external-memory output path (north star: "writes into the engine's image via CUDA-Vulkan external-memory interop").

No Vulkan loader exists in this image, so a VkDeviceMemory fd cannot be produced here.  What CAN be produced is the other kind of
exportable device memory: a CUDA virtual-memory allocation (cuMemCreate with a POSIX-fd shareable handle).  The test hands that fd
to mm_bind_output_external_buffer_fd -- the same cudaImportExternalMemory(OPAQUE_FD) -> cudaExternalMemoryGetMappedBuffer sequence a
Vulkan-exported linear image goes through -- and records the driver's answer:
  * accepted: the march renders through the imported mapping and the pixels, read back through the EXPORTER's own mapping of the
    same physical memory, equal a plain render bit for bit;
  * refused: the refusal must be a clean MM_ERR_CUDA with the driver's message (printed), and the context stays usable.
The opaque-fd array path (mm_bind_output_external_fd -> mipmapped array -> surface) and the semaphore imports take a Vulkan-side
object that cannot be made here; their argument and error behaviour is what is tested.
"""
import os

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def _vmm_export(nbytes, device=0):
    """-> (fd, exporter's device pointer, padded size, cleanup()) for a POSIX-fd-exportable CUDA VMM allocation"""
    from cuda.bindings import driver as cu

    def ok(res):
        err = res[0]
        if err != cu.CUresult.CUDA_SUCCESS:
            raise RuntimeError(f"{err}")
        return res[1] if len(res) == 2 else res[1:]
    ok(cu.cuInit(0))
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = device
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = ok(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    size = (nbytes + gran - 1) // gran * gran
    handle = ok(cu.cuMemCreate(size, prop, 0))
    fd = ok(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0))
    va = ok(cu.cuMemAddressReserve(size, 0, 0, 0))
    ok(cu.cuMemMap(va, size, 0, handle, 0))
    acc = cu.CUmemAccessDesc()
    acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = device
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ok(cu.cuMemSetAccess(va, size, [acc], 1))

    def cleanup():
        cu.cuMemUnmap(va, size)
        cu.cuMemAddressFree(va, size)
        cu.cuMemRelease(handle)
    return int(fd), int(va), size, cleanup


def test_external_linear_memory_import_answer_is_recorded(mm, assets):
    import torch
    torch.cuda.init()
    torch.cuda.synchronize()
    W, H = 160, 90
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.allocOutput()
    plain = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
    try:
        fd, va, size, cleanup = _vmm_export(W * H * 16)
    except Exception as e:                                   # the exporter side is test scaffolding, not the product
        cs.close()
        pytest.skip(f"cannot create an exportable VMM allocation here: {e}")
    try:
        try:
            cs.bindOutputExternalBufferFd(fd, size)
            accepted = True
        except mm.MarshmallowError as e:
            accepted = False
            print("DRIVER ANSWER: cudaImportExternalMemory(OPAQUE_FD) of a CUDA-VMM fd was REFUSED:", e)
            assert e.code == -2 and "cuda" in str(e).lower()
            os.close(fd)
        if accepted:
            print("DRIVER ANSWER: cudaImportExternalMemory(OPAQUE_FD) of a CUDA-VMM fd was ACCEPTED; rendering through the imported mapping")
            cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
            cs.dispatch()
            cs.synchronize()
            view = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
            from cuda.bindings import runtime as rt
            err, = rt.cudaMemcpy(view.data_ptr(), va, W * H * 16, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
            assert int(err) == 0
            got = view.cpu().numpy()
            assert np.array_equal(got.view(np.uint32), plain.view(np.uint32)), "pixels written through the imported mapping differ"
            assert np.array_equal(cs.readOutput().view(np.uint32), plain.view(np.uint32))
        # either way the context is still usable with its own image afterwards
        cs.allocOutput()
        again = cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
        assert np.array_equal(again.view(np.uint32), plain.view(np.uint32))
    finally:
        cs.close()
        cleanup()


def test_external_imports_reject_what_is_not_an_exported_object(mm, assets):
    """Argument and error behaviour of the entries that need a Vulkan-side object: no crash, a status and a message."""
    cs = mm.ComputeShader(0, (64, 36))
    lib = mm.load_library()
    import ctypes as C
    r, w = os.pipe()                                         # a valid fd that is no exported memory / semaphore object
    try:
        assert lib.mm_bind_output_external_fd(cs._ctx, -1, 64 * 36 * 16, 64, 36) == -1
        assert lib.mm_bind_output_external_fd(cs._ctx, r, 16, 64, 36) == -1                      # allocation smaller than the image
        assert lib.mm_bind_output_external_buffer_fd(cs._ctx, r, 64 * 36 * 16, 8, 64 * 16, 64, 36) == -1    # misaligned offset
        assert lib.mm_bind_output_external_buffer_fd(cs._ctx, r, 64 * 36 * 16, 0, 64 * 8, 64, 36) == -1     # pitch < 16*w
        rc = lib.mm_bind_output_external_buffer_fd(cs._ctx, r, 64 * 36 * 16, 0, 64 * 16, 64, 36)
        assert rc == -2, rc
        print("pipe fd as external memory ->", lib.mm_last_error(cs._ctx).decode())
        slot = C.c_int(-1)
        assert lib.mm_import_semaphore_fd(cs._ctx, -1, 0, C.byref(slot)) == -1
        rc = lib.mm_import_semaphore_fd(cs._ctx, w, 0, C.byref(slot))
        assert rc == -2, rc
        print("pipe fd as external semaphore ->", lib.mm_last_error(cs._ctx).decode())
        for fn in (lib.mm_signal_semaphore, lib.mm_wait_semaphore):
            assert fn(cs._ctx, 0, 0, None) == -3                                                   # nothing imported in slot 0
            assert fn(cs._ctx, 9, 0, None) == -3
        assert lib.mm_release_semaphore(cs._ctx, 0) == -3
        assert lib.mm_enable_peer(cs._ctx, cs._ctx) == 0                                           # same device: nothing to do
    finally:
        for f in (r, w):
            try:
                os.close(f)
            except OSError:
                pass
        cs.close()
