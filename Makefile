# Builds the C-ABI CUDA library (sm_100a only) in-tree so the .so travels with the repo snapshot.
NVCC      ?= nvcc
PKG       := semantic_pyramid_for_image_generation_b200
CSRC      := $(PKG)/csrc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
LIB       := $(PKG)/libspyramid_b200.so

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/conv_halo_common.cuh include/spyramid_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

native_tests: $(LIB) build/test_conv_native

# halo_probe.cu is a developer probe (descriptor validation of the halo-tiled operand): test-only, never in the product .so
build/test_conv_native: tests/native/test_conv_native.cu tests/native/halo_probe.cu $(LIB)
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ tests/native/test_conv_native.cu tests/native/halo_probe.cu -L$(PKG) -lspyramid_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)'

# developer probes behind the numbers in profiles/r01_mma_probe.txt and r01_tma_probe.txt
probes: $(LIB) build/mma_probe build/tma_probe build/pdl_probe

build/pdl_probe: tests/native/pdl_probe.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ $<

build/mma_probe: tests/native/mma_probe.cu $(CSRC)/common.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ $<

build/tma_probe: tests/native/tma_probe.cu $(LIB)
	$(NVCC) $(ARCH) -O2 -std=c++17 -o $@ $< -L$(PKG) -lspyramid_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)'

clean:
	rm -rf build $(LIB)

.PHONY: all native_tests probes clean
