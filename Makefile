# Native build without Python (the same commands j3d_b200/build.py runs).
#   make            libj3dg.so (nvcc, sm_100a only) + libj3dg_host.so (g++) + libj3d_synth.so (gcc)
#   make oracle     the test oracles (oracle/Makefile): plain-C restatement, and the unmodified reference if J3D_REF exists
#   make shim       tests/cpp/shim_frame, the C++ driver of the scene / canvas mirror (needs a B200 to run)
NVCC ?= nvcc
CXX ?= g++
CC ?= gcc
PKG := j3d_b200
NVCC_FLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-fvisibility=hidden --expt-relaxed-constexpr

all: $(PKG)/libj3dg.so $(PKG)/libj3dg_host.so $(PKG)/libj3d_synth.so

$(PKG)/libj3dg.so: $(wildcard $(PKG)/csrc/*.cu) $(wildcard $(PKG)/csrc/*.cuh) include/j3dg.h
	$(NVCC) $(NVCC_FLAGS) -shared -I include -I $(PKG)/csrc -o $@ $(wildcard $(PKG)/csrc/*.cu)

$(PKG)/libj3dg_host.so: $(wildcard $(PKG)/host/*.cpp) $(wildcard $(PKG)/host/*.h) include/j3dg.h
	$(CXX) -std=c++17 -O2 -ffp-contract=off -fPIC -shared -Wall -I include -I $(PKG)/host -o $@ $(wildcard $(PKG)/host/*.cpp) -ldl

$(PKG)/libj3d_synth.so: $(PKG)/synth/synth.c
	$(CC) -std=c11 -O2 -ffp-contract=off -fopenmp -fPIC -shared -Wall -o $@ $< -lm

oracle:
	$(MAKE) -C oracle oracle
	@if [ -f "$${J3D_REF:-/root/reference}/j3d/canvas.cpp" ]; then $(MAKE) -C oracle -j8 ref J3D_REF=$${J3D_REF:-/root/reference}; fi

shim: all
	mkdir -p build/tests
	$(CXX) -std=c++17 -O1 -ffp-contract=off -Wall -Werror -I include -I $(PKG)/host tests/cpp/shim_frame.cpp -o build/tests/shim_frame \
	  -L $(PKG) -lj3dg -lj3dg_host -Wl,-rpath,$(abspath $(PKG))

clean:
	rm -f $(PKG)/*.so build/tests/shim_frame

.PHONY: all oracle shim clean
