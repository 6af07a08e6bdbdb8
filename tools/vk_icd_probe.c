// tools/vk_icd_probe.c -- is there a usable Vulkan implementation on this box WITHOUT a loader?
// The GPU boxes ship no libvulkan.so.1 and no ICD manifest, but the NVIDIA user-mode driver libraries that contain
// the Vulkan ICD (libGLX_nvidia.so.0 / libEGL_nvidia.so.0) may be present.  An ICD exports vk_icdGetInstanceProcAddr
// (loader-ICD interface); calling it directly works without a loader.  Hand-declared prototypes: no Vulkan headers
// exist in this image.  Build: gcc -O1 -o tools/vk_icd_probe tools/vk_icd_probe.c -ldl
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef void (*PFN_vkVoidFunction)(void);
typedef PFN_vkVoidFunction (*PFN_GetInstanceProcAddr)(void* instance, const char* name);
typedef struct { int sType; const void* pNext; const char* pApplicationName; uint32_t applicationVersion;
                 const char* pEngineName; uint32_t engineVersion; uint32_t apiVersion; } AppInfo;
typedef struct { int sType; const void* pNext; uint32_t flags; const AppInfo* pApplicationInfo; uint32_t enabledLayerCount;
                 const char* const* ppEnabledLayerNames; uint32_t enabledExtensionCount;
                 const char* const* ppEnabledExtensionNames; } InstanceCreateInfo;
typedef struct { char extensionName[256]; uint32_t specVersion; } ExtensionProperties;

int main(void)
{
  const char* libs[] = {"libGLX_nvidia.so.0", "libEGL_nvidia.so.0", "libvulkan.so.1", "libnvidia-glcore.so", 0};
  for(int i = 0; libs[i]; ++i)
  {
    void* h = dlopen(libs[i], RTLD_NOW | RTLD_LOCAL);
    printf("== %s: %s\n", libs[i], h ? "loaded" : dlerror());
    if(!h)
      continue;
    PFN_GetInstanceProcAddr gipa = (PFN_GetInstanceProcAddr)dlsym(h, "vk_icdGetInstanceProcAddr");
    if(!gipa)
      gipa = (PFN_GetInstanceProcAddr)dlsym(h, "vkGetInstanceProcAddr");
    printf("   vk_icdGetInstanceProcAddr: %p\n", (void*)gipa);
    if(!gipa)
      continue;
    int (*negotiate)(uint32_t*) = (int (*)(uint32_t*))dlsym(h, "vk_icdNegotiateLoaderICDInterfaceVersion");
    if(negotiate)
    {
      uint32_t v = 5;
      int      r = negotiate(&v);
      printf("   negotiate -> %d, interface version %u\n", r, v);
    }
    int (*createInstance)(const InstanceCreateInfo*, const void*, void**) =
        (int (*)(const InstanceCreateInfo*, const void*, void**))gipa(0, "vkCreateInstance");
    printf("   vkCreateInstance: %p\n", (void*)createInstance);
    if(!createInstance)
      continue;
    AppInfo            app = {0, 0, "nvpyr-probe", 1, "none", 1, (1u << 22) | (1u << 12)};  // VK_API_VERSION_1_1
    InstanceCreateInfo ci  = {1, 0, 0, &app, 0, 0, 0, 0};
    void*              inst = 0;
    int                r    = createInstance(&ci, 0, &inst);
    printf("   vkCreateInstance -> %d, instance %p\n", r, inst);
    fflush(stdout);
    if(r != 0 || !inst)
      continue;
    int (*enumPhys)(void*, uint32_t*, void**) = (int (*)(void*, uint32_t*, void**))gipa(inst, "vkEnumeratePhysicalDevices");
    void (*getProps)(void*, void*)            = (void (*)(void*, void*))gipa(inst, "vkGetPhysicalDeviceProperties");
    int (*enumExt)(void*, const char*, uint32_t*, ExtensionProperties*) =
        (int (*)(void*, const char*, uint32_t*, ExtensionProperties*))gipa(inst, "vkEnumerateDeviceExtensionProperties");
    uint32_t n = 0;
    void*    phys[16];
    r = enumPhys(inst, &n, 0);
    printf("   vkEnumeratePhysicalDevices -> %d, count %u\n", r, n);
    fflush(stdout);
    if(n > 16)
      n = 16;
    if(n == 0)
      continue;
    enumPhys(inst, &n, phys);
    for(uint32_t d = 0; d < n; ++d)
    {
      static unsigned char props[4096];
      memset(props, 0, sizeof props);
      getProps(phys[d], props);
      uint32_t api, drv, vendor;
      memcpy(&api, props, 4), memcpy(&drv, props + 4, 4), memcpy(&vendor, props + 8, 4);
      printf("   device %u: '%s' api %u.%u.%u vendor 0x%x driver 0x%x\n", d, (const char*)props + 20, api >> 22, (api >> 12) & 1023,
             api & 4095, vendor, drv);
      uint32_t ne = 0;
      enumExt(phys[d], 0, &ne, 0);
      static ExtensionProperties ext[1024];
      if(ne > 1024)
        ne = 1024;
      enumExt(phys[d], 0, &ne, ext);
      printf("   %u device extensions; of interest:", ne);
      for(uint32_t e = 0; e < ne; ++e)
        if(strstr(ext[e].extensionName, "external_memory") || strstr(ext[e].extensionName, "glsl_shader")
           || strstr(ext[e].extensionName, "external_semaphore") || strstr(ext[e].extensionName, "timeline")
           || strstr(ext[e].extensionName, "subgroup"))
          printf(" %s", ext[e].extensionName);
      printf("\n");
    }
    return 0;
  }
  printf("== no Vulkan ICD could create an instance\n");
  return 1;
}
