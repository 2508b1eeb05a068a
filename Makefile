# Builds, in-tree:
#   jpeg_rust_b200/lib/libjpgpu.so   the product: C ABI + sm_100a kernels (nvcc)
#   jpeg_rust_b200/lib/libjpgenc.so  offline input generator (g++)
#   oracle/liboracle.so              CPU oracle — test infrastructure only (gcc)
#   tests/sim/libjpsim.so            CPU simulation of the parallel algorithm — tests only (g++)
#   tests/c/abi_smoke                plain C program linking libjpgpu.so — tests only (gcc)
NVCC ?= nvcc
CXX ?= g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
CSRC = jpeg_rust_b200/csrc
LIB = jpeg_rust_b200/lib

all: $(LIB)/libjpgpu.so $(LIB)/libjpgenc.so oracle tests/sim/libjpsim.so tests/c/abi_smoke

$(LIB)/libjpgpu.so: $(CSRC)/jpgpu_kernels.cu $(CSRC)/jpgpu_api.cu $(CSRC)/jpgpu_host.cpp $(CSRC)/jpgpu_multi.cpp $(CSRC)/jpgpu_core.h $(CSRC)/jpgpu_kernels.cuh $(CSRC)/jpgpu_host.h include/jpgpu.h
	@mkdir -p $(LIB)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/jpgpu_kernels.cu $(CSRC)/jpgpu_api.cu $(CSRC)/jpgpu_host.cpp $(CSRC)/jpgpu_multi.cpp -lcudart -lpthread 2> $(LIB)/ptxas_report.txt || (cat $(LIB)/ptxas_report.txt; false)
	@grep -v "Compile time" $(LIB)/ptxas_report.txt > $(LIB)/ptxas_report.tmp; mv $(LIB)/ptxas_report.tmp $(LIB)/ptxas_report.txt
	@grep -E "error|warning" $(LIB)/ptxas_report.txt | grep -v "ptxas info" || true

$(LIB)/libjpgenc.so: $(CSRC)/jpgenc.cpp
	@mkdir -p $(LIB)
	$(CXX) -O2 -fPIC -shared -std=c++17 -Wall -o $@ $<

oracle:
	$(MAKE) -s -C oracle

tests/sim/libjpsim.so: tests/sim/jpsim.cpp $(CSRC)/jpgpu_host.cpp $(CSRC)/jpgpu_core.h $(CSRC)/jpgpu_host.h
	$(CXX) -O2 -fPIC -shared -std=c++17 -Wall -I/usr/local/cuda/include -o $@ tests/sim/jpsim.cpp $(CSRC)/jpgpu_host.cpp

# plain C consumer of the ABI (gcc + include/jpgpu.h + the shared library, nothing else)
tests/c/abi_smoke: tests/c/abi_smoke.c include/jpgpu.h $(LIB)/libjpgpu.so
	gcc -std=c11 -O1 -Wall -Wextra -Iinclude -o $@ tests/c/abi_smoke.c -L$(LIB) -ljpgpu -Wl,-rpath,'$$ORIGIN/../../$(LIB)'

clean:
	rm -f tests/c/abi_smoke $(LIB)/*.so $(LIB)/ptxas_report.txt tests/sim/*.so oracle/*.so

.PHONY: all oracle clean
