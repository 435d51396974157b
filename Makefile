# Build the B200-native engine (libpgm_b200.so) and the C step oracle (oracle/liboracle_step.so).
NVCC ?= nvcc
CXX ?= g++
CC ?= gcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xfatbin=-compress-all -Xcompiler -fPIC,-O3,-Wall
CSRC := pogema_b200/csrc
BUILD := build
LIB := pogema_b200/_lib/libpgm_b200.so
INST := step_priority step_block_both step_soft observe reset
FAST := priority block_both soft
CU := pgm_capi pgm_plan pgm_transport pgm_devgen $(foreach i,$(INST),pgm_inst_$(i)_g0 pgm_inst_$(i)_g1) $(foreach i,$(FAST),pgm_inst_fast_$(i)_a pgm_inst_fast_$(i)_b)
OBJS := $(addprefix $(BUILD)/,$(addsuffix .o,$(CU))) $(BUILD)/pgm_gen.o $(BUILD)/pgm_hostexpand.o
HDRS := $(CSRC)/pgm_devgen.h $(CSRC)/pgm_kernels.cuh $(CSRC)/pgm_fast.cuh $(CSRC)/pgm_fast_launch.cuh $(CSRC)/pgm_launch.cuh $(CSRC)/pgm_engine.h $(CSRC)/pgm_rng.h $(CSRC)/pgm_gen.h $(CSRC)/pgm_hostexpand.h include/pgm_b200.h

all: $(LIB) oracle

$(BUILD)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) $(PTXAS_V) -c $< -o $@

$(BUILD)/pgm_gen.o: $(CSRC)/pgm_gen.cpp $(HDRS)
	@mkdir -p $(BUILD)
	$(CXX) -O3 -std=c++17 -fPIC -Wall -c $< -o $@

$(BUILD)/pgm_hostexpand.o: $(CSRC)/pgm_hostexpand.cpp $(HDRS)
	@mkdir -p $(BUILD)
	$(CXX) -O3 -std=c++17 -fPIC -Wall -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p pogema_b200/_lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lpthread

# Timeline build: the same library with the phase stamps compiled in (tools/phase_timeline.py loads it through
# PGM_B200_LIB).  Not part of `all`: the product build carries no stamp code.
TL_BUILD := build_tl
TL_LIB := pogema_b200/_lib/libpgm_b200_timeline.so
TL_OBJS := $(addprefix $(TL_BUILD)/,$(addsuffix .o,$(CU))) $(BUILD)/pgm_gen.o $(BUILD)/pgm_hostexpand.o

$(TL_BUILD)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(TL_BUILD)
	$(NVCC) $(NVFLAGS) -DPGM_TIMELINE -c $< -o $@

$(TL_LIB): $(TL_OBJS)
	@mkdir -p pogema_b200/_lib
	$(NVCC) $(ARCH) -shared -o $@ $(TL_OBJS) -lpthread

timeline: $(TL_LIB)

oracle: oracle/liboracle_step.so

oracle/liboracle_step.so: oracle/step_oracle.c
	$(CC) -O2 -fPIC -shared -Wall -o $@ $< -lm

clean:
	rm -rf $(BUILD) $(TL_BUILD) $(LIB) $(TL_LIB) oracle/liboracle_step.so

.PHONY: all oracle clean timeline
