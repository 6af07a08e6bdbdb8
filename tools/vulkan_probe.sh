#!/bin/bash
# Probe the GPU box for any Vulkan implementation (loader, ICDs, lavapipe, tools).  Output is committed
# under profiles/ as the evidence for SURVEY §8f rank 1 / the north_star's lavapipe baseline.
echo "== date: $(date -u)"; echo "== uname: $(uname -a)"
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,driver_version,pci.bus_id --format=csv
echo "== ldconfig -p | grep -i -E 'vulkan|lvp|GLX_nvidia|nvidia-glcore|nvidia-gpucomp'"; ldconfig -p | grep -i -E 'vulkan|lvp|GLX_nvidia|nvidia-glcore|nvidia-gpucomp|nvidia-glvkspirv'
echo "== find libvulkan / ICD json / lavapipe"
find / -xdev \( -name 'libvulkan*' -o -name '*_icd*.json' -o -name 'nvidia_icd*.json' -o -name 'libvulkan_lvp*' -o -name 'lvp_icd*' -o -name 'libGLX_nvidia*' -o -name 'libnvidia-glvkspirv*' -o -name 'libnvidia-gpucomp*' -o -name 'libnvidia-glcore*' -o -name 'libEGL_nvidia*' \) 2>/dev/null | head -50
for d in /usr/share/vulkan /etc/vulkan /usr/local/share/vulkan /usr/share/glvnd /etc/glvnd; do echo "== ls -R $d"; ls -R $d 2>&1 | head -20; done
echo "== tools"; for t in vulkaninfo glslangValidator glslc spirv-as spirv-val; do printf "%s: " $t; command -v $t || echo absent; done
echo "== libnvidia-* present"; ls /usr/lib/x86_64-linux-gnu/ 2>/dev/null | grep -i nvidia | head -60
echo "== NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
echo "== python dlopen probe"
python - <<'PY'
import ctypes
for n in ("libvulkan.so.1", "libvulkan.so", "libGLX_nvidia.so.0", "libEGL_nvidia.so.0", "libnvidia-vulkan-producer.so", "libvulkan_lvp.so"):
    try:
        ctypes.CDLL(n); print(n, "LOADS")
    except OSError as e:
        print(n, "absent:", str(e)[:100])
PY
