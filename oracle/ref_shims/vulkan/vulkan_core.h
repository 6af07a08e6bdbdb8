// Mock of the handful of Vulkan declarations that the reference's
// nvpro_pyramid/nvpro_pyramid_dispatch.hpp uses.  The vkCmd* entry points
// append to a per-thread event log (defined in oracle/ref_harness.cpp) so the
// unmodified reference scheduler can be run on the CPU and its recorded
// command sequence compared with our planner.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <stdint.h>
#include <stddef.h>

typedef struct MockPipeline_T*       VkPipeline;
typedef struct MockPipelineLayout_T* VkPipelineLayout;
typedef struct MockCommandBuffer_T*  VkCommandBuffer;
#define VK_NULL_HANDLE nullptr

typedef uint32_t VkFlags;
typedef VkFlags  VkAccessFlags;
typedef VkFlags  VkPipelineStageFlags;
typedef VkFlags  VkShaderStageFlags;
typedef VkFlags  VkDependencyFlags;
typedef enum { VK_STRUCTURE_TYPE_MEMORY_BARRIER = 46 } VkStructureType;
typedef enum { VK_PIPELINE_BIND_POINT_COMPUTE = 1 } VkPipelineBindPoint;
enum { VK_ACCESS_SHADER_READ_BIT = 0x20, VK_ACCESS_SHADER_WRITE_BIT = 0x40 };
enum { VK_PIPELINE_STAGE_COMPUTE_SHADER_BIT = 0x800 };
enum { VK_SHADER_STAGE_COMPUTE_BIT = 0x20 };

typedef struct VkMemoryBarrier
{
  VkStructureType sType;
  const void*     pNext;
  VkAccessFlags   srcAccessMask;
  VkAccessFlags   dstAccessMask;
} VkMemoryBarrier;
struct VkBufferMemoryBarrier;
struct VkImageMemoryBarrier;

void vkCmdBindPipeline(VkCommandBuffer, VkPipelineBindPoint, VkPipeline);
void vkCmdPushConstants(VkCommandBuffer, VkPipelineLayout, VkShaderStageFlags, uint32_t offset, uint32_t size,
                        const void* pValues);
void vkCmdDispatch(VkCommandBuffer, uint32_t x, uint32_t y, uint32_t z);
void vkCmdPipelineBarrier(VkCommandBuffer, VkPipelineStageFlags, VkPipelineStageFlags, VkDependencyFlags,
                          uint32_t memoryBarrierCount, const VkMemoryBarrier*, uint32_t,
                          const VkBufferMemoryBarrier*, uint32_t, const VkImageMemoryBarrier*);
