# Native build: libautognothi_b200.so (CUDA kernels + C-ABI, sm_100a only) and the native checks.
# `python -c "import __graft_entry__ as g; g.build()"` drives the same rules.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr
CSRC      := autognothi_b200/csrc
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
LIB       := autognothi_b200/lib/libautognothi_b200.so
TLIB      := autognothi_b200/lib/libagb_torch.so
NATIVE    := build/gemm_check
PYTHON    ?= python
TORCH_DIR := $(shell $(PYTHON) -c "import torch,os;print(os.path.dirname(torch.__file__))")
TORCH_ABI := $(shell $(PYTHON) -c "import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")

all: $(LIB) $(TLIB) $(NATIVE) oracle

# thin torch C++ extension over the C-ABI: one TORCH_LIBRARY op per entry point, generated from the header
build/agb_torch_binding.cpp: include/autognothi_b200.h tools/gen_torch_binding.py
	@mkdir -p build
	$(PYTHON) tools/gen_torch_binding.py include/autognothi_b200.h > $@

$(TLIB): build/agb_torch_binding.cpp $(LIB)
	g++ -O2 -std=c++17 -fPIC -shared -D_GLIBCXX_USE_CXX11_ABI=$(TORCH_ABI) -Iinclude -I$(TORCH_DIR)/include \
	    -I$(TORCH_DIR)/include/torch/csrc/api/include -I/usr/local/cuda/include $< -o $@ \
	    -L$(TORCH_DIR)/lib -ltorch -ltorch_cpu -lc10 -lc10_cuda -Lautognothi_b200/lib -lautognothi_b200 -Wl,-rpath,'$$ORIGIN'

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
	rm -rf build $(LIB) $(TLIB)

.PHONY: all clean oracle
