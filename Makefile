# Build of the three shared libraries and the headless executable (all in-tree, git-ignored, shipped to the GPU box by gpurun):
#   vulkan-restir-pt_b200/lib/librestirpt.so        CUDA kernels + C ABI (include/restirpt.h), sm_100a only
#   vulkan-restir-pt_b200/lib/librestirpt_host.so   C++ host (Scene / Camera / Renderer), include/restirpt_host.h
#   vulkan-restir-pt_b200/bin/restirpt_render       the reference's main.cpp, headless (scene -> frames -> PNG)
#   oracle/liboracle.so                             CPU oracle — TEST INFRASTRUCTURE, never linked by the three above
PKG := vulkan-restir-pt_b200
LIBDIR ?= $(PKG)/lib
# experiment builds: make cuda host LIBDIR=<dir> BUILD=<dir> EXTRA=-D<macro>; loaded with RPT_LIB_DIR=<dir>
BUILD ?= build
EXTRA ?=
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++

# Numeric contract: no implicit FMA contraction on either side, so the CUDA kernels and
# the oracle produce bit-identical fp32 results; fused multiply-adds are spelled out in the sources.
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
             -ftz=false -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Iinclude -diag-suppress 20012 $(EXTRA)
HOSTFLAGS := -O2 -std=c++17 -fPIC -Wall -Iinclude
ORCFLAGS := -O2 -std=c++17 -fPIC -Wall -ffp-contract=off -mfma -Iinclude -pthread

CUDA_SRCS := $(wildcard $(PKG)/csrc/*.cu)
CUDA_HDRS := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/restirpt.h
HOST_SRCS := $(filter-out $(PKG)/host/main.cpp,$(wildcard $(PKG)/host/*.cpp))
HOST_HDRS := $(wildcard $(PKG)/host/*.h) include/restirpt.h include/restirpt_host.h
ORC_SRCS := $(wildcard oracle/*.cpp)
ORC_HDRS := $(wildcard oracle/*.h) include/restirpt.h

all: cuda host oracle cli

cuda: $(LIBDIR)/librestirpt.so
host: $(LIBDIR)/librestirpt_host.so
oracle: oracle/liboracle.so
cli: $(PKG)/bin/restirpt_render

CUDA_OBJS := $(patsubst $(PKG)/csrc/%.cu,$(BUILD)/%.o,$(CUDA_SRCS))

$(BUILD)/%.o: $(PKG)/csrc/%.cu $(CUDA_HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVCCFLAGS) -c -o $@ $<

$(LIBDIR)/librestirpt.so: $(CUDA_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(CUDA_OBJS)

$(LIBDIR)/librestirpt_host.so: $(HOST_SRCS) $(HOST_HDRS) $(LIBDIR)/librestirpt.so
	@mkdir -p $(LIBDIR)
	$(CXX) $(HOSTFLAGS) -shared -o $@ $(HOST_SRCS) -L$(LIBDIR) -lrestirpt -Wl,-rpath,'$$ORIGIN'

# the headless executable (reference src/main.cpp): host library + CUDA library, found next to it through the rpath
$(PKG)/bin/restirpt_render: $(PKG)/host/main.cpp $(HOST_HDRS) $(LIBDIR)/librestirpt_host.so
	@mkdir -p $(PKG)/bin
	$(CXX) $(HOSTFLAGS) -o $@ $(PKG)/host/main.cpp -L$(LIBDIR) -lrestirpt_host -lrestirpt -Wl,-rpath,'$$ORIGIN/../lib'

oracle/liboracle.so: $(ORC_SRCS) $(ORC_HDRS)
	$(CXX) $(ORCFLAGS) -shared -o $@ $(ORC_SRCS)

clean:
	rm -f $(LIBDIR)/*.so oracle/liboracle.so $(PKG)/bin/restirpt_render

.PHONY: all cuda host oracle cli clean
