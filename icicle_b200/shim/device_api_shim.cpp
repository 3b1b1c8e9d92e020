// libicicle_backend_cuda_device.so : DeviceAPI for the B200 backend, registered under B200_DEVICE_TYPE ("CUDA").
// Implements every pure virtual of icicle::DeviceAPI (icicle/include/icicle/device_api.h:44-182) by forwarding to the
// C ABI (include/icicle_b200.h).  Model: icicle/backend/cpu/src/cpu_device_api.cpp:9-105.
// The file name contains "device" so that the loader dlopens it RTLD_GLOBAL (icicle/src/runtime.cpp:306-323).
#include "shim_common.h"
#include "icicle/device_api.h"

using namespace icicle;
using b200_shim::to_err;

class B200DeviceAPI : public DeviceAPI
{
public:
  eIcicleError set_device(const Device& device) override { return to_err(b200_set_device(device.id)); }
  eIcicleError get_device_count(int& device_count) const override { return to_err(b200_get_device_count(&device_count)); }

  eIcicleError allocate_memory(void** ptr, size_t size) const override { return to_err(b200_malloc(ptr, size)); }
  eIcicleError allocate_memory_async(void** ptr, size_t size, icicleStreamHandle stream) const override
  {
    return to_err(b200_malloc_async(ptr, size, stream));
  }
  eIcicleError free_memory(void* ptr) const override { return to_err(b200_free(ptr)); }
  eIcicleError free_memory_async(void* ptr, icicleStreamHandle stream) const override { return to_err(b200_free_async(ptr, stream)); }
  eIcicleError get_available_memory(size_t& total, size_t& free) const override
  {
    return to_err(b200_get_available_memory(&total, &free));
  }
  eIcicleError memset(void* ptr, int value, size_t size) const override { return to_err(b200_memset(ptr, value, size)); }
  eIcicleError memset_async(void* ptr, int value, size_t size, icicleStreamHandle stream) const override
  {
    return to_err(b200_memset_async(ptr, value, size, stream));
  }

  eIcicleError copy(void* dst, const void* src, size_t size, eCopyDirection direction) const override
  {
    return do_copy(dst, src, size, direction, nullptr, 0);
  }
  eIcicleError copy_async(void* dst, const void* src, size_t size, eCopyDirection direction, icicleStreamHandle stream) const override
  {
    return do_copy(dst, src, size, direction, stream, 1);
  }

  eIcicleError synchronize(icicleStreamHandle stream = nullptr) const override { return to_err(b200_synchronize(stream)); }
  eIcicleError create_stream(icicleStreamHandle* stream) const override { return to_err(b200_create_stream(stream)); }
  eIcicleError destroy_stream(icicleStreamHandle stream) const override { return to_err(b200_destroy_stream(stream)); }

  eIcicleError get_device_properties(DeviceProperties& properties) const override
  {
    properties.using_host_memory = false;
    properties.num_memory_regions = 0;
    properties.supports_pinned_memory = true;
    return eIcicleError::SUCCESS;
  }

private:
  static eIcicleError do_copy(void* dst, const void* src, size_t size, eCopyDirection direction, icicleStreamHandle stream, int is_async)
  {
    switch (direction) {
    case eCopyDirection::HostToDevice: return to_err(b200_copy_to_device(dst, src, size, stream, is_async));
    case eCopyDirection::DeviceToHost: return to_err(b200_copy_to_host(dst, src, size, stream, is_async));
    case eCopyDirection::DeviceToDevice: return to_err(b200_copy_device_to_device(dst, src, size, stream, is_async));
    case eCopyDirection::HostToHost: std::memcpy(dst, src, size); return eIcicleError::SUCCESS;
    }
    return eIcicleError::INVALID_ARGUMENT;
  }
};

REGISTER_DEVICE_API(B200_DEVICE_TYPE, B200DeviceAPI);
