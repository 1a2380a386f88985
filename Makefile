# Build of the B200-native RNA-SeQC hot path.
#   make            -> product: CUDA library + host helper library + CLI     (needs nvcc; cross-compiles sm_100a)
#   make oracle     -> test infrastructure: CPU oracle library + oracle-backed CLI harness
#   make ref        -> test infrastructure: the unmodified reference binary into oracle/_ref/ (needs /root/reference)
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      := /usr/bin/g++
CXXFLAGS := -std=c++17 -O2 -fPIC -Wall -Wno-sign-compare -Iinclude -Irnaseqc_b200/csrc/host
NVFLAGS  := -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Iinclude
LIBDIR   := rnaseqc_b200/lib
BINDIR   := rnaseqc_b200/bin
OBJDIR   := build
ODIR     := oracle/_build

HOST_SRCS := annotation.cpp alnreader.cpp cramreader.cpp bamparallel.cpp fastinflate.cpp fasta.cpp report.cpp pipeline.cpp
HOST_OBJS := $(addprefix $(OBJDIR)/host/,$(HOST_SRCS:.cpp=.o))

all: $(LIBDIR)/librsq_b200.so $(LIBDIR)/librsq_host.so $(BINDIR)/rnaseqc
oracle: $(ODIR)/librsq_oracle.so $(ODIR)/rnaseqc_oracle $(ODIR)/libhostcore.so $(ODIR)/libhostcore_sd.so $(ODIR)/libhostlegacy.so $(ODIR)/librsq_emu.so $(ODIR)/rnaseqc_emu
ref:
	$(MAKE) -f oracle/Makefile.ref

$(OBJDIR)/host/%.o: rnaseqc_b200/csrc/host/%.cpp $(wildcard rnaseqc_b200/csrc/host/*.h) include/rsq_b200.h
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@

CU_SRCS := rnaseqc_b200/csrc/rsq_kernels.cu rnaseqc_b200/csrc/rsq_covstats.cu rnaseqc_b200/csrc/rsq_capi.cu
$(LIBDIR)/librsq_b200.so: $(CU_SRCS) rnaseqc_b200/csrc/rsq_device.h rnaseqc_b200/csrc/rsq_legacy.cuh rnaseqc_b200/csrc/rsq_hot_core.cuh rnaseqc_b200/csrc/rsq_eit.h include/rsq_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o $@ $(CU_SRCS) -lcudart 2> $(LIBDIR)/ptxas.log || (cat $(LIBDIR)/ptxas.log; false)

$(LIBDIR)/librsq_host.so: $(HOST_OBJS) $(OBJDIR)/host/host_capi.o
	@mkdir -p $(LIBDIR)
	$(CXX) -shared -o $@ $^ -lz -ldl -lpthread

$(BINDIR)/rnaseqc: $(HOST_OBJS) $(OBJDIR)/host/cli_main.o $(LIBDIR)/librsq_b200.so
	@mkdir -p $(BINDIR)
	$(CXX) -o $@ $(HOST_OBJS) $(OBJDIR)/host/cli_main.o -L$(LIBDIR) -lrsq_b200 -lz -lpthread -Wl,-rpath,'$$ORIGIN/../lib'

$(ODIR)/librsq_oracle.so: oracle/rsq_oracle.cpp include/rsq_b200.h
	@mkdir -p $(ODIR)
	$(CXX) -std=c++17 -O2 -fPIC -Wall -Wno-sign-compare -Iinclude -shared -o $@ $<

# test infrastructure: host emulation of the hot kernel's per-record logic + index builder (tests/test_hostcore.py)
$(ODIR)/libhostcore.so: tests/hostcore/hostcore.cpp rnaseqc_b200/csrc/rsq_hot_core.cuh rnaseqc_b200/csrc/rsq_eit.h rnaseqc_b200/csrc/rsq_device.h include/rsq_b200.h
	@mkdir -p $(ODIR)
	$(CXX) -std=c++17 -O2 -fPIC -Wall -Wno-sign-compare -Wno-unused-function -Iinclude -Irnaseqc_b200/csrc -I/usr/local/cuda/include -x c++ -shared -o $@ $<
# host emulation of the --legacy device rules (rsq_legacy.cuh compiled for the CPU behind shims)
$(ODIR)/libhostlegacy.so: tests/hostcore/hostlegacy.cpp rnaseqc_b200/csrc/rsq_legacy.cuh rnaseqc_b200/csrc/rsq_hot_core.cuh rnaseqc_b200/csrc/rsq_device.h include/rsq_b200.h
	@mkdir -p $(ODIR)
	$(CXX) -std=c++17 -O2 -fPIC -Wall -Wno-sign-compare -Wno-unused-function -Iinclude -Irnaseqc_b200/csrc -I/usr/local/cuda/include -x c++ -shared -o $@ $<
# the same with the kernel's experimental compile-time variants that change the per-record logic (tools/ab_build.sh builds the CUDA side)
$(ODIR)/libhostcore_sd.so: tests/hostcore/hostcore.cpp rnaseqc_b200/csrc/rsq_hot_core.cuh rnaseqc_b200/csrc/rsq_eit.h rnaseqc_b200/csrc/rsq_device.h include/rsq_b200.h
	@mkdir -p $(ODIR)
	$(CXX) -std=c++17 -O2 -fPIC -Wall -Wno-sign-compare -Wno-unused-function -DRSQ_STAGE_DESC -DRSQ_STAGE_N=64 -Iinclude -Irnaseqc_b200/csrc -I/usr/local/cuda/include -x c++ -shared -o $@ $<

# test infrastructure: the WHOLE CUDA library compiled for the host through tests/cudaemu/cuda_emu.h (kernel and C-ABI sources unchanged; one
# OS thread per CUDA thread; barriers, warp collectives, atomics and the runtime calls emulated): librsq_emu.so implements include/rsq_b200.h
# on the CPU, rnaseqc_emu is the product CLI over it — tests/test_cudaemu.py, and the -m gpu suite replayed with RSQ_EMULATE=1
EMU_SRCS := tests/cudaemu/emu_kernels.cpp tests/cudaemu/emu_covstats.cpp tests/cudaemu/emu_capi.cpp
EMU_DEPS := tests/cudaemu/cuda_emu.h $(CU_SRCS) rnaseqc_b200/csrc/rsq_device.h rnaseqc_b200/csrc/rsq_legacy.cuh rnaseqc_b200/csrc/rsq_hot_core.cuh rnaseqc_b200/csrc/rsq_eit.h include/rsq_b200.h
EMU_FLAGS := -std=c++17 -O2 -fPIC -Wall -Wno-sign-compare -Wno-unused-function -Wno-unused-value -Wno-attributes -Wno-unknown-pragmas -Iinclude -Irnaseqc_b200/csrc -Itests/cudaemu -I/usr/local/cuda/include
$(ODIR)/emu_%.o: tests/cudaemu/emu_%.cpp $(EMU_DEPS)
	@mkdir -p $(ODIR)
	$(CXX) $(EMU_FLAGS) -c $< -o $@
$(ODIR)/librsq_emu.so: $(ODIR)/emu_kernels.o $(ODIR)/emu_covstats.o $(ODIR)/emu_capi.o
	$(CXX) -shared -o $@ $^ -lpthread -ldl
$(ODIR)/rnaseqc_emu: $(HOST_OBJS) $(OBJDIR)/host/cli_main.o $(ODIR)/librsq_emu.so
	$(CXX) -o $@ $(HOST_OBJS) $(OBJDIR)/host/cli_main.o -L$(ODIR) -lrsq_emu -lz -lpthread -Wl,-rpath,'$$ORIGIN'

$(ODIR)/rnaseqc_oracle: $(HOST_OBJS) $(OBJDIR)/host/cli_main.o $(ODIR)/librsq_oracle.so
	$(CXX) -o $@ $(HOST_OBJS) $(OBJDIR)/host/cli_main.o -L$(ODIR) -lrsq_oracle -lz -lpthread -Wl,-rpath,'$$ORIGIN'

# host-side measurement tool: BAM ingest records/s by thread count (RSQ_INGEST_STATS=1 prints the busy time of every stage)
tools: $(OBJDIR)/decode_bench $(OBJDIR)/inflate_bench
# zlib against the ingest's own DEFLATE decoder on the BGZF blocks of a BAM, one thread
$(OBJDIR)/inflate_bench: tools/inflate_bench.cpp rnaseqc_b200/csrc/host/fastinflate.cpp rnaseqc_b200/csrc/host/fastinflate.h
	@mkdir -p $(OBJDIR)
	$(CXX) -std=c++17 -O2 -Irnaseqc_b200/csrc/host -o $@ tools/inflate_bench.cpp rnaseqc_b200/csrc/host/fastinflate.cpp -lz
$(OBJDIR)/decode_bench: tools/decode_bench.cpp $(HOST_OBJS)
	$(CXX) $(CXXFLAGS) -o $@ tools/decode_bench.cpp $(filter-out $(OBJDIR)/host/report.o $(OBJDIR)/host/pipeline.o,$(HOST_OBJS)) -lz -lpthread

# the host side (CLI over the CPU oracle backend) under AddressSanitizer + UBSan: build/rnaseqc_asan — run it over BAMs by hand or put it
# in the place of oracle/_build/rnaseqc_oracle in tools/ingest_mutation_fuzz.py (ASAN_OPTIONS=detect_leaks=0:exitcode=99 RSQ_FULL_TEARDOWN=1)
asan:
	@mkdir -p $(OBJDIR)/asan
	for f in $(HOST_SRCS:.cpp=) cli_main; do $(CXX) -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-sign-compare -Iinclude -Irnaseqc_b200/csrc/host -c rnaseqc_b200/csrc/host/$$f.cpp -o $(OBJDIR)/asan/$$f.o || exit 1; done
	$(CXX) -std=c++17 -O1 -g -fsanitize=address,undefined -Wno-sign-compare -Iinclude -c oracle/rsq_oracle.cpp -o $(OBJDIR)/asan/oracle.o
	$(CXX) -fsanitize=address,undefined -o $(OBJDIR)/rnaseqc_asan $(OBJDIR)/asan/*.o -lz -lpthread

clean:
	rm -rf $(OBJDIR) $(LIBDIR) $(BINDIR) $(ODIR)
.PHONY: all oracle ref tools asan clean
