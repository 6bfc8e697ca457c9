# Builds libdwt_b200/libdwtb200.so (CUDA kernels + C ABI, sm_100a only) and the checkers under oracle/.
NVCC ?= nvcc
ARCH  = -gencode arch=compute_100a,code=sm_100a
# DEBUG_KEYS=1 compiles the measurement-only tuning keys 96-99 and the no-store / no-arithmetic branches of k_fwd_level in
NVFLAGS = $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v $(if $(DEBUG_KEYS),-DDWTB200_DEBUG_KEYS)
CSRC = libdwt_b200/csrc
OBJS = $(CSRC)/kernels_stream.o $(CSRC)/kernels_ring.o $(CSRC)/kernels_ring2.o $(CSRC)/kernels_tail.o $(CSRC)/kernels_tile.o $(CSRC)/kernels_inplace.o $(CSRC)/kernels_vol.o $(CSRC)/kernels_generic.o $(CSRC)/kernels_util.o $(CSRC)/strips.o $(CSRC)/dwtb200.o

all: libdwt_b200/libdwtb200.so libdwt_b200/libdwt_compat.so oracle examples

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/dwtb200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(CSRC)/$*.ptxas.log || (cat $(CSRC)/$*.ptxas.log; false)

libdwt_b200/libdwtb200.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared -lrt

libdwt_b200/libdwt_compat.so: $(CSRC)/libdwt_compat.c include/libdwt_compat.h include/dwtb200.h libdwt_b200/libdwtb200.so
	gcc -std=c99 -O2 -fPIC -Wall -Wextra -shared -o $@ $< -Llibdwt_b200 -ldwtb200 -Wl,-rpath,'$$ORIGIN'

# The reference's UNMODIFIED example programs, compiled from where they lie and linked against the
# B200 library first and the compiled reference (for everything outside the hot path) second.
# Build container only (needs $(REF)); the binaries travel to the GPU box in build/ (git-ignored).
# examples/test (the reference's own round-trip test program) calls the transforms only through the reference's
# dwt_util_test2_* helpers, so the compat library is forced in front (--no-as-needed) for its entry points to be the ones they reach.
REF ?= /root/reference
EXAMPLES = simple simple-int simple-double simple-perf simple-perf-int simple-single-loop simple-perf-single
examples: libdwt_b200/libdwt_compat.so oracle
	@if [ -d $(REF)/examples ]; then mkdir -p build/examples; for e in $(EXAMPLES); do \
	  src=$(REF)/examples/$$e/simple.c; \
	  gcc -std=c99 -O2 -D_POSIX_C_SOURCE=199309L -D_GNU_SOURCE -I$(REF)/src $$src -o build/examples/$$e \
	    -Llibdwt_b200 -ldwt_compat -ldwtb200 -Loracle/_ref -l:libdwt_ref.so -lm -lrt -fopenmp \
	    -Wl,-rpath,'$$ORIGIN/../../libdwt_b200:$$ORIGIN/../../oracle/_ref' || exit 1; done; \
	  for e in test/test subbands/subbands subbands-int/subbands load/simple load-int/simple simple-newapi/simple; do \
	  gcc -std=c99 -O2 -D_POSIX_C_SOURCE=199309L -D_GNU_SOURCE -I$(REF)/src $(REF)/examples/$$e.c -o build/examples/$${e%%/*} \
	    -Llibdwt_b200 -Wl,--no-as-needed -ldwt_compat -ldwtb200 -Wl,--as-needed -Loracle/_ref -l:libdwt_ref.so -lm -lrt -fopenmp \
	    -Wl,-rpath,'$$ORIGIN/../../libdwt_b200:$$ORIGIN/../../oracle/_ref' || exit 1; done; \
	 else echo "examples: $(REF) absent, keeping prebuilt build/examples (if any)"; fi

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.ptxas.log libdwt_b200/libdwtb200.so libdwt_b200/libdwt_compat.so

.PHONY: all oracle clean examples
