# Builds libdwt_b200/libdwtb200.so (CUDA kernels + C ABI, sm_100a only) and the checkers under oracle/.
NVCC ?= nvcc
ARCH  = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v
CSRC = libdwt_b200/csrc
OBJS = $(CSRC)/kernels_stream.o $(CSRC)/kernels_tail.o $(CSRC)/kernels_generic.o $(CSRC)/kernels_util.o $(CSRC)/dwtb200.o

all: libdwt_b200/libdwtb200.so oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/kernels.h $(CSRC)/lifting.cuh include/dwtb200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(CSRC)/$*.ptxas.log || (cat $(CSRC)/$*.ptxas.log; false)

libdwt_b200/libdwtb200.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.ptxas.log libdwt_b200/libdwtb200.so

.PHONY: all oracle clean
