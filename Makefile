# libdwt_b200.so (CUDA sm_100a + C host layer), the encode/decode CLIs, and the test oracle.
NVCC    ?= /usr/local/cuda/bin/nvcc
CC      ?= gcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Idwt_b200/csrc --use_fast_math $(EXTRA)
CFLAGS  := -std=c99 -O2 -W -Wall -fPIC -Iinclude
CSRC    := dwt_b200/csrc
LIB     := dwt_b200/libdwt_b200.so
CU      := $(CSRC)/lift.cu $(CSRC)/hilbert.cu $(CSRC)/coder_enc.cu $(CSRC)/coder_dec.cu $(CSRC)/pipeline.cu \
           $(CSRC)/pipeline_dec.cu $(CSRC)/capi.cu
OBJ     := $(CU:.cu=.o) dwt_b200/host/streamio.o
HDR     := $(wildcard $(CSRC)/*.cuh) include/dwt_b200.h dwt_b200/host/streamio_internal.h

all: $(LIB) encode decode dwtbatch oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(HDR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $@.ptxas.log || (cat $@.ptxas.log; false)

dwt_b200/host/%.o: dwt_b200/host/%.c $(HDR)
	$(CC) $(CFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -cudart static

# drop-in CLIs: same argv, exit codes and stderr lines as the reference programs
encode: dwt_b200/host/encode.c dwt_b200/host/pnm.c $(LIB)
	$(CC) $(CFLAGS) dwt_b200/host/encode.c dwt_b200/host/pnm.c -o $@ -Ldwt_b200 -ldwt_b200 -Wl,-rpath,'$$ORIGIN/dwt_b200'
decode: dwt_b200/host/decode.c dwt_b200/host/pnm.c $(LIB)
	$(CC) $(CFLAGS) dwt_b200/host/decode.c dwt_b200/host/pnm.c -o $@ -Ldwt_b200 -ldwt_b200 -Wl,-rpath,'$$ORIGIN/dwt_b200'
# many images per process on a pool of contexts (no reference counterpart: SURVEY.md 8f-3)
dwtbatch: dwt_b200/host/batch.c dwt_b200/host/pnm.c $(LIB)
	$(CC) $(CFLAGS) dwt_b200/host/batch.c dwt_b200/host/pnm.c -o $@ -Ldwt_b200 -ldwt_b200 -Wl,-rpath,'$$ORIGIN/dwt_b200'

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(OBJ) $(CSRC)/*.ptxas.log $(LIB) encode decode dwtbatch
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
