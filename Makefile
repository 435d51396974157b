# Build the B200-native engine (libpgm_b200.so) and the C step oracle (oracle/liboracle_step.so).
NVCC ?= nvcc
CXX ?= g++
CC ?= gcc
ARCH := -gencode arch=compute_100a,code=sm_100a
CSRC := pogema_b200/csrc
LIB := pogema_b200/_lib/libpgm_b200.so

all: $(LIB) oracle

$(LIB): $(CSRC)/pgm_capi.cu $(CSRC)/pgm_kernels.cuh $(CSRC)/pgm_rng.h $(CSRC)/pgm_gen.cpp $(CSRC)/pgm_gen.h include/pgm_b200.h
	mkdir -p pogema_b200/_lib
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Xptxas -v -Xcompiler -fPIC,-O3,-Wall -shared \
	    $(CSRC)/pgm_capi.cu $(CSRC)/pgm_gen.cpp -o $(LIB) -lpthread

oracle: oracle/liboracle_step.so

oracle/liboracle_step.so: oracle/step_oracle.c
	$(CC) -O2 -fPIC -shared -Wall -o $@ $< -lm

clean:
	rm -f $(LIB) oracle/liboracle_step.so

.PHONY: all oracle clean
