#!/bin/bash
# can the reference's own oracle (Nim + Mesa llvmpipe + Xvfb) run on the GPU box?
# Prints what it finds; the answer is recorded in BASELINE.md.
echo "== toolchain"
for t in nim nimble atlas choosenim Xvfb xvfb-run glxinfo eglinfo clang glslangValidator; do
  p=$(command -v $t 2>/dev/null); echo "$t: ${p:-MISSING}"
done
echo "== GL / EGL / OSMesa / X11 libraries known to the loader"
ldconfig -p 2>/dev/null | grep -i -e "libGL\." -e libOSMesa -e libEGL -e libX11 -e libGLX -e libgallium -e swrast || echo "none"
echo "== Mesa DRI drivers on disk"
find / -xdev \( -name "swrast_dri.so" -o -name "libgallium*.so" -o -name "libOSMesa*" -o -name "*llvmpipe*" \) 2>/dev/null | head -20
echo "== any libGL on disk"
find / -xdev -name "libGL.so*" 2>/dev/null | head -20
echo "== host"
nproc; grep -m1 "model name" /proc/cpuinfo
