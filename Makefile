# nsparse-b200 build.  Targets keep the names of the reference cuda-c/Makefile where they overlap
# (spgemm_hash_s/d, amb_s/d); the sample drivers are compiled UNCHANGED from the reference tree when
# it is present (REF_DIR), against include/nsparse.h and libnsparse_{s,d}.a.
NVCC      ?= nvcc
CXX       := /usr/bin/g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
# make EXTRA=-DNSP_PHASE_TIMING rebuilds the heavy numeric kernel with per-phase cycle counters
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fopenmp --expt-extended-lambda -Iinclude $(EXTRA)
REF_DIR   ?= /root/reference
SRC       := nsparse_b200/csrc
OBJ       := build/obj
LIBDIR    := nsparse_b200/lib
BIN       := bin

CORE_CU   := context peer_push peer_dma spgemm_plan spgemm_symbolic spgemm_numeric_s spgemm_numeric_d c_api mgpu_api amb_convert amb_spmv amb_api
CORE_OBJ  := $(addprefix $(OBJ)/,$(addsuffix .o,$(CORE_CU))) $(OBJ)/gen.o $(OBJ)/mtx_reader.o

.PHONY: all lib compat drivers clean oracle
all: lib compat oracle

lib: $(LIBDIR)/libnsparse_b200.so $(LIBDIR)/libnsparse_gen.so

$(OBJ)/%.o: $(SRC)/%.cu $(wildcard $(SRC)/*.h $(SRC)/*.cuh include/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/gen.o: $(SRC)/gen.cpp include/nsparse_b200.h
	@mkdir -p $(OBJ)
	$(CXX) -O3 -fPIC -fopenmp -Iinclude -c $< -o $@

$(OBJ)/mtx_reader.o: $(SRC)/mtx_reader.cpp include/nsparse_b200.h
	@mkdir -p $(OBJ)
	$(CXX) -O3 -std=c++17 -fPIC -fopenmp -Iinclude -c $< -o $@

$(LIBDIR)/libnsparse_b200.so: $(CORE_OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $^ -Xcompiler -fopenmp -lcudart -lgomp

# the synthetic-input generator on its own (plain C++, no CUDA): bench.py's reference arm and the tests generate
# their inputs without loading the product library
$(LIBDIR)/libnsparse_gen.so: $(OBJ)/gen.o
	@mkdir -p $(LIBDIR)
	$(CXX) -shared -fopenmp -o $@ $^

# ---- the nsparse.h API, one archive per precision (reference: .s.o / .d.o objects) ----
compat: $(LIBDIR)/libnsparse_s.a $(LIBDIR)/libnsparse_d.a

$(OBJ)/compat_api.s.o: $(SRC)/compat_api.cu $(wildcard $(SRC)/*.h $(SRC)/*.cuh include/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DFLOAT -c $< -o $@
$(OBJ)/compat_api.d.o: $(SRC)/compat_api.cu $(wildcard $(SRC)/*.h $(SRC)/*.cuh include/*.h)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DDOUBLE -c $< -o $@
$(OBJ)/compat_cusparse.s.o: $(SRC)/compat_cusparse.cu include/nsparse.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DFLOAT -c $< -o $@
$(OBJ)/compat_cusparse.d.o: $(SRC)/compat_cusparse.cu include/nsparse.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DDOUBLE -c $< -o $@

$(LIBDIR)/libnsparse_s.a: $(OBJ)/compat_api.s.o $(OBJ)/compat_cusparse.s.o $(CORE_OBJ)
	@mkdir -p $(LIBDIR)
	rm -f $@ && ar rcs $@ $^
$(LIBDIR)/libnsparse_d.a: $(OBJ)/compat_api.d.o $(OBJ)/compat_cusparse.d.o $(CORE_OBJ)
	@mkdir -p $(LIBDIR)
	rm -f $@ && ar rcs $@ $^

# ---- unchanged reference sample drivers (only when the reference tree is mounted) ----
DRV_FLAGS := $(ARCH) -O3 -Iinclude -Wno-deprecated-declarations -Xcompiler -fopenmp
drivers: compat
	@mkdir -p $(BIN)
	@if [ -d $(REF_DIR)/cuda-c/src/sample ]; then \
	  set -e; \
	  $(NVCC) $(DRV_FLAGS) -DFLOAT  $(REF_DIR)/cuda-c/src/sample/spgemm/spgemm_hash.cu -o $(BIN)/spgemm_hash_s -L$(LIBDIR) -lnsparse_s -lcusparse -lgomp; \
	  $(NVCC) $(DRV_FLAGS) -DDOUBLE $(REF_DIR)/cuda-c/src/sample/spgemm/spgemm_hash.cu -o $(BIN)/spgemm_hash_d -L$(LIBDIR) -lnsparse_d -lcusparse -lgomp; \
	  $(NVCC) $(DRV_FLAGS) -DFLOAT  $(REF_DIR)/cuda-c/src/sample/spmv/spmv_amb.cu -o $(BIN)/amb_s -L$(LIBDIR) -lnsparse_s -lcusparse -lgomp; \
	  $(NVCC) $(DRV_FLAGS) -DDOUBLE $(REF_DIR)/cuda-c/src/sample/spmv/spmv_amb.cu -o $(BIN)/amb_d -L$(LIBDIR) -lnsparse_d -lcusparse -lgomp; \
	  $(NVCC) $(DRV_FLAGS) -DFLOAT  $(REF_DIR)/cuda-c/src/sample/spgemm/spgemm_cu_csr.cu -o $(BIN)/spgemm_cu_csr_s -L$(LIBDIR) -lnsparse_s -lcusparse -lgomp; \
	  $(NVCC) $(DRV_FLAGS) -DDOUBLE $(REF_DIR)/cuda-c/src/sample/spgemm/spgemm_cu_csr.cu -o $(BIN)/spgemm_cu_csr_d -L$(LIBDIR) -lnsparse_d -lcusparse -lgomp; \
	  echo "drivers built from $(REF_DIR) (sources unchanged)"; \
	else echo "reference tree not present: drivers skipped"; fi
	$(NVCC) $(DRV_FLAGS) -DFLOAT  $(SRC)/sample/spmv_cu_csr.cu -o $(BIN)/cu_csr_s -L$(LIBDIR) -lnsparse_s -lcusparse -lgomp
	$(NVCC) $(DRV_FLAGS) -DDOUBLE $(SRC)/sample/spmv_cu_csr.cu -o $(BIN)/cu_csr_d -L$(LIBDIR) -lnsparse_d -lcusparse -lgomp
	$(NVCC) $(DRV_FLAGS) -DFLOAT  $(SRC)/sample/spgemm_hash_mgpu.cu -o $(BIN)/spgemm_hash_mgpu_s -L$(LIBDIR) -lnsparse_s -lcusparse -lgomp
	$(NVCC) $(DRV_FLAGS) -DDOUBLE $(SRC)/sample/spgemm_hash_mgpu.cu -o $(BIN)/spgemm_hash_mgpu_d -L$(LIBDIR) -lnsparse_d -lcusparse -lgomp
	@# the same protocol driver that oracle/Makefile builds against the REFERENCE's SpGEMM, linked against this
	@# library instead (test infrastructure: raw CSR in, reference timing protocol, C out)
	@mkdir -p oracle/_ref
	$(NVCC) $(DRV_FLAGS) -DNSP_OURS -DFLOAT  oracle/ref_gpu/dump_spgemm.cu -o oracle/_ref/dump_spgemm_ours_s -L$(LIBDIR) -lnsparse_s -lcusparse -lgomp
	$(NVCC) $(DRV_FLAGS) -DNSP_OURS -DDOUBLE oracle/ref_gpu/dump_spgemm.cu -o oracle/_ref/dump_spgemm_ours_d -L$(LIBDIR) -lnsparse_d -lcusparse -lgomp

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIBDIR) $(BIN)
	$(MAKE) -C oracle clean
