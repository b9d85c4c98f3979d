# Native build: libautognothi_b200.so (CUDA kernels + C-ABI, sm_100a only) and the native checks.
# `python -c "import __graft_entry__ as g; g.build()"` drives the same rules.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr
CSRC      := autognothi_b200/csrc
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
LIB       := autognothi_b200/lib/libautognothi_b200.so
NATIVE    := build/gemm_check

all: $(LIB) $(NATIVE) oracle

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/autognothi_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p autognothi_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

build/%: tests/native/%.cu $(LIB) include/autognothi_b200.h
	$(NVCC) $(ARCH) -O2 -std=c++17 $< -o $@ -Lautognothi_b200/lib -lautognothi_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../autognothi_b200/lib'

oracle:
	@if [ -f oracle/Makefile ]; then $(MAKE) -C oracle; fi

clean:
	rm -rf build $(LIB)

.PHONY: all clean oracle
