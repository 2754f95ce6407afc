# Builds libb200pt.so (CUDA kernels + C ABI + headless host) and the b200pt CLI for sm_100a.
# `python -c "import __graft_entry__ as g; g.build()"` drives this file.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
PKG       := rtx-pathtracer_b200
BUILD     := $(PKG)/_build
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v --expt-relaxed-constexpr
CXXFLAGS  := -O2 -std=c++17 -fPIC -ffp-contract=off -Wall

LIB       := $(PKG)/libb200pt.so
CLI       := $(PKG)/b200pt
CU_SRCS   := $(PKG)/csrc/api.cu $(PKG)/csrc/guiding_fit.cu
HOST_SRCS := $(PKG)/host/scene.cpp $(PKG)/host/bvh.cpp $(PKG)/host/exr.cpp $(PKG)/host/app.cpp $(PKG)/host/image_decode.cpp
CU_OBJS   := $(patsubst $(PKG)/csrc/%.cu,$(BUILD)/%.o,$(CU_SRCS))
HOST_OBJS := $(patsubst $(PKG)/host/%.cpp,$(BUILD)/%.o,$(HOST_SRCS))
HEADERS   := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/host/*.h) include/b200pt.h include/b200pt_detmath.h

all: $(LIB) $(CLI)

# per-file extras (GUIDING_FLAGS: tuning overrides for guiding_fit.cu, e.g. -DG_BLOCK=256 -DG_BLOCKS_PER_SM=2)
GUIDING_FLAGS ?=
EXTRA_guiding_fit := $(GUIDING_FLAGS)
TRACER_FLAGS ?=
EXTRA_api := $(TRACER_FLAGS)

$(BUILD)/%.o: $(PKG)/csrc/%.cu $(HEADERS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) $(EXTRA_$*) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(BUILD)/%.o: $(PKG)/host/%.cpp $(HEADERS)
	@mkdir -p $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(CU_OBJS) $(HOST_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lz -cudart static

$(CLI): $(PKG)/host/main.cpp $(LIB) include/b200pt.h
	$(CXX) $(CXXFLAGS) -o $@ $< -L$(PKG) -lb200pt -Wl,-rpath,'$$ORIGIN'

clean:
	rm -rf $(BUILD) $(LIB) $(CLI)

.PHONY: all clean
