# Builds the C-ABI library of the B200 ShadowKV decode path in-tree (the .so ships to the GPU box with the
# gpurun snapshot).  sm_100a only.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
ifdef TRACE
NVCCFLAGS += -DSKV_TRACE
endif
SRC := $(wildcard shadowkv_b200/csrc/*.cu)
BUILD ?= build
OBJ := $(patsubst shadowkv_b200/csrc/%.cu,$(BUILD)/%.o,$(SRC))
LIB ?= shadowkv_b200/lib/libshadowkv_b200.so

all: $(LIB)

$(BUILD)/%.o: shadowkv_b200/csrc/%.cu $(wildcard shadowkv_b200/csrc/*.cuh) include/shadowkv_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p shadowkv_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -lcudart

# phase-timeline build for kernel tuning (tools/trace_kernels.py); never loaded by the product path
trace:
	$(MAKE) TRACE=1 BUILD=build_trace LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so

clean:
	rm -rf build build_trace $(LIB) shadowkv_b200/lib/libshadowkv_b200_trace.so
.PHONY: all clean trace
