# Builds the B200-native library, the `raft` CLI on top of it, and the test oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
SRC       := raft_b200/csrc
OBJ       := build
CU        := nametable k0_fasta k1_paf k2_coverage k3_repeat_cut k5_emit api
OBJS      := $(addprefix $(OBJ)/,$(addsuffix .o,$(CU))) $(OBJ)/host_io.o
LIB       := raft_b200/libraft_b200.so

all: $(LIB) raft_b200/raft raft_b200/split_naive raft_b200/libraft_synth.so oracle

$(OBJ)/%.o: $(SRC)/%.cu $(SRC)/common.cuh $(SRC)/kernels.h $(SRC)/nametable.cuh $(SRC)/coverage.cuh include/raft_b200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/host_io.o: $(SRC)/host_io.cpp $(SRC)/file_io.h include/raft_b200.h
	@mkdir -p $(OBJ)
	$(CXX) -O2 -std=c++17 -fPIC -Wall -I/usr/local/cuda/include -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lz -lnccl

raft_b200/libraft_synth.so: $(SRC)/synth_gen.cu $(SRC)/common.cuh
	$(NVCC) $(NVFLAGS) -shared $< -o $@

raft_b200/raft: $(SRC)/raft_main.cpp $(LIB)
	$(CXX) -O2 -std=c++17 -Wall $< -o $@ -Lraft_b200 -lraft_b200 -Wl,-rpath,'$$ORIGIN'

raft_b200/split_naive: $(SRC)/split_naive_main.cpp $(LIB)
	$(CXX) -O2 -std=c++17 -Wall $< -o $@ -Lraft_b200 -lraft_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(OBJ) $(LIB) raft_b200/raft raft_b200/split_naive raft_b200/libraft_synth.so
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
