# Builds, in-tree:
#   raytrace_b200/lib/librt_b200.so   CUDA kernels + C ABI (include/rt_b200.h), sm_100a only
#   raytrace_b200/lib/librt_host.so   C++ object model (Scene/Model/RayTracer...) + C shim for ctypes
#   raytrace_b200/bin/rt_render       headless driver (tools/render_main.cpp)
#   oracle/liboracle.so               TEST INFRASTRUCTURE: CPU restatement of the reference path
#   oracle/_ref/ref_render            TEST INFRASTRUCTURE: the reference itself (needs /root/reference)
NVCC ?= nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the parity path needs one rounding per operation (see csrc/rt_device.cuh)
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off -Xptxas -v
HOSTFLAGS := -std=c++17 -O2 -fPIC -ffp-contract=off -Wall -Wno-unused-function

P := raytrace_b200
CSRC := $(P)/csrc/rt_api.cu $(P)/csrc/rt_kernels.cu $(P)/csrc/rt_build.cu
CHDR := $(wildcard $(P)/csrc/*.cuh $(P)/csrc/*.h) include/rt_b200.h
HSRC := $(wildcard $(P)/host/*.cpp)
HHDR := $(wildcard $(P)/host/*.h) $(P)/scenes/scenes.h include/rt_b200.h

all: $(P)/lib/librt_b200.so $(P)/lib/librt_host.so $(P)/bin/rt_render oracle/liboracle.so ref

$(P)/lib/%.o: $(P)/csrc/%.cu $(CHDR)
	@mkdir -p $(P)/lib
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

$(P)/lib/librt_b200.so: $(P)/lib/rt_api.o $(P)/lib/rt_kernels.o $(P)/lib/rt_build.o
	$(NVCC) $(ARCH) -shared -o $@ $^

$(P)/lib/librt_host.so: $(HSRC) $(HHDR) $(P)/lib/librt_b200.so
	$(CXX) $(HOSTFLAGS) -shared -o $@ $(HSRC) -L$(P)/lib -lrt_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

$(P)/bin/rt_render: tools/render_main.cpp $(P)/lib/librt_host.so $(HHDR)
	@mkdir -p $(P)/bin
	$(CXX) $(HOSTFLAGS) -I$(P)/host -o $@ tools/render_main.cpp -L$(P)/lib -lrt_host -lrt_b200 -Wl,-rpath,'$$ORIGIN/../lib' -lpthread

oracle/liboracle.so: oracle/rt_oracle.cpp include/rt_b200.h
	$(CXX) $(HOSTFLAGS) -shared -o $@ oracle/rt_oracle.cpp -lpthread

ref:
	@oracle/build_ref.sh

clean:
	rm -rf $(P)/lib $(P)/bin oracle/liboracle.so

.PHONY: all ref clean

# A/B builds during development: `make variant NAME=n256 VFLAGS="-DRT_NODE_FETCH=1"` builds
# raytrace_b200/lib_n256/{librt_b200.so,librt_host.so}; select it with RT_B200_LIBDIR=raytrace_b200/lib_n256
variant:
	@mkdir -p $(P)/lib_$(NAME)
	for f in rt_api rt_kernels rt_build; do $(NVCC) $(NVFLAGS) $(VFLAGS) -c $(P)/csrc/$$f.cu -o $(P)/lib_$(NAME)/$$f.o 2> $(P)/lib_$(NAME)/$$f.o.ptxas.log || { cat $(P)/lib_$(NAME)/$$f.o.ptxas.log; exit 1; }; done
	$(NVCC) $(ARCH) -shared -o $(P)/lib_$(NAME)/librt_b200.so $(P)/lib_$(NAME)/rt_api.o $(P)/lib_$(NAME)/rt_kernels.o $(P)/lib_$(NAME)/rt_build.o
	$(CXX) $(HOSTFLAGS) -shared -o $(P)/lib_$(NAME)/librt_host.so $(HSRC) -L$(P)/lib_$(NAME) -lrt_b200 -Wl,-rpath,'$$ORIGIN' -lpthread
.PHONY: variant
