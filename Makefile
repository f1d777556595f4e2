# Builds libsfsim_atmosphere.so, the C-ABI CUDA library behind `clj -T:build atmosphere-lut`.
# Mirrors the reference's rule for its only native library (sfsim Makefile:1-18: one shared object
# in the repository root, loaded with (ffi/load-library ...)).
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH = -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the double-precision index maps and geometry must round like the reference's JVM
# arithmetic; the hot loops ask for their FMAs explicitly (fmaf / fma).
NVCCFLAGS = $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -Wall -Iinclude
CSRC = sfsim_b200/csrc
SRCS = $(CSRC)/atm_api.cu $(CSRC)/atm_tables.cu $(CSRC)/atm_lookup.cu $(CSRC)/atm_batch.cu $(CSRC)/noise.cu $(CSRC)/cubemap.cu
HDRS = $(CSRC)/atm_math.cuh $(CSRC)/atm_device.cuh $(CSRC)/atm_kernel_common.cuh $(CSRC)/atm_tables.h $(CSRC)/atm_api_internal.h include/sfsim_atmosphere.h include/sfsim_noise.h include/sfsim_cubemap.h
OBJS = $(SRCS:.cu=.o)

all: libsfsim_atmosphere.so

%.o: %.cu $(HDRS)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

libsfsim_atmosphere.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log libsfsim_atmosphere.so

.PHONY: all oracle clean
