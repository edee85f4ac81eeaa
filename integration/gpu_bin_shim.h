// gpu_bin_shim.h -- compiles the reference's BinModule.cpp, unmodified, with the GPU binding in place.
//
// oracle/Makefile (target ref_gpu) passes this file with `-include` when it compiles
// /root/reference/fastore/fastore_bin/BinModule.cpp.  It first includes every header BinModule.cpp includes, so
// that all the reference's declarations are seen with their real names (the include guards make the later
// #includes no-ops), then defines stand-ins with the interface of FastqCategorizerSE/PE and
// FastqRecordsPackerSE/PE whose Categorize + PackToBins pair goes through GpuBinEncoder, and finally renames the
// four class names for the body of BinModule.cpp only.  The two lines at BinModule.cpp:130-133 and :379-382
// thereby become one fsb_bin_chunks call; nothing of the reference is copied or edited.
//
// Scope: the -t1 loops of BinModuleSE/PE::Fastq2Bin.  The multi-thread operators (BinOperator.cpp) work on the bins
// map between the two calls (small-bin buffering) and stay on the CPU.
#ifndef H_GPU_BIN_SHIM
#define H_GPU_BIN_SHIM

#include "Globals.h"

#include <vector>
#include <array>
#include <iostream>

#include "BinModule.h"
#include "FastqStream.h"
#include "FastqParser.h"
#include "FastqPacker.h"
#include "FastqCategorizer.h"
#include "BinFile.h"
#include "BinOperator.h"
#include "Exception.h"
#include "Thread.h"

#include "GpuBinEncoder.h"

namespace gpu_shim {
// Categorize hands the parsed reads to PackToBins of the same thread (the reference calls them back to back)
inline std::vector<FastqRecord>*& pending() { static thread_local std::vector<FastqRecord>* p = NULL; return p; }
}

template <class RealCategorizer>
class GpuCategorizerT : public RealCategorizer
{
public:
    GpuCategorizerT(const MinimizerParameters& params_, const MinimizerFilteringParameters& filter_ = MinimizerFilteringParameters(),
                    const CategorizerParameters& catParams_ = CategorizerParameters())
        : RealCategorizer(params_, filter_, catParams_) {}
    void Categorize(std::vector<FastqRecord>& records_, std::map<uint32, FastqRecordsPtrBin>& bins_)
    {
        bins_.clear();                          // the bins map stays empty: binning happens on the device, inside PackToBins
        gpu_shim::pending() = &records_;
    }
};

template <class RealPacker>
class GpuPackerT : public RealPacker            // UnpackFromBin (fastore_bin d) stays the reference's
{
public:
    GpuPackerT(const BinModuleConfig& binConfig_) : RealPacker(binConfig_), gpu(binConfig_) {}
    void PackToBins(const std::map<uint32, FastqRecordsPtrBin>& /*dnaBins_*/, BinaryBinBlock& binBlock_)
    {
        std::vector<FastqRecord>* reads = gpu_shim::pending();
        if (reads == NULL) throw Exception("GpuPacker: PackToBins without Categorize");
        gpu.CategorizeAndPack(*reads, binBlock_);
        gpu_shim::pending() = NULL;
    }
private:
    GpuBinEncoder gpu;
};

typedef GpuCategorizerT<FastqCategorizerSE> GpuCategorizerSE;
typedef GpuCategorizerT<FastqCategorizerPE> GpuCategorizerPE;
typedef GpuPackerT<FastqRecordsPackerSE> GpuPackerSE;
typedef GpuPackerT<FastqRecordsPackerPE> GpuPackerPE;

#define FastqCategorizerSE GpuCategorizerSE
#define FastqCategorizerPE GpuCategorizerPE
#define FastqRecordsPackerSE GpuPackerSE
#define FastqRecordsPackerPE GpuPackerPE

#endif // H_GPU_BIN_SHIM
