# Builds the C-ABI library of the B200 ShadowKV decode path in-tree (the .so ships to the GPU box with the
# gpurun snapshot).  sm_100a only.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
SRC := $(wildcard shadowkv_b200/csrc/*.cu)
OBJ := $(patsubst shadowkv_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := shadowkv_b200/lib/libshadowkv_b200.so

all: $(LIB)

build/%.o: shadowkv_b200/csrc/%.cu $(wildcard shadowkv_b200/csrc/*.cuh) include/shadowkv_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p shadowkv_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -lcudart

clean:
	rm -rf build $(LIB)
.PHONY: all clean
