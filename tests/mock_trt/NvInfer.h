// MOCK of the TensorRT 8.x plugin interfaces -- TEST INFRASTRUCTURE, not TensorRT.
//
// This image has no TensorRT.  This header declares, from the public TensorRT 8.6
// API documentation, exactly the types and virtuals that (a) include/trtx_plugins.h (our header-only
// adapters) and (b) the reference's own plugin sources need in order to COMPILE, plus a tiny in-process
// plugin registry so REGISTER_TENSORRT_PLUGIN / getPluginRegistry()->getPluginCreator() work in tests.
// It contains no TensorRT code.  When a real <NvInfer.h> is on the include path it must be used instead.
#ifndef TRTX_MOCK_NVINFER_H
#define TRTX_MOCK_NVINFER_H

#include <cuda_runtime_api.h>

#include <cassert>
#include <cfloat>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <utility>

// The reference's rcnn plugins override some virtuals WITHOUT noexcept (rcnn/BatchedNmsPlugin.h:46), i.e.
// they only compile against the TensorRT 7 flavour of the interfaces; everything else targets TensorRT 8.
// -DTRTX_MOCK_TRT_MAJOR=7 selects the pre-noexcept / non-const-enqueue flavour.
#ifndef TRTX_MOCK_TRT_MAJOR
#define TRTX_MOCK_TRT_MAJOR 8
#endif
#define NV_TENSORRT_MAJOR TRTX_MOCK_TRT_MAJOR
#define NV_TENSORRT_MINOR 6
#define NV_TENSORRT_PATCH 1
#define NV_TENSORRT_VERSION (TRTX_MOCK_TRT_MAJOR * 1000 + 601)
#define TRTX_MOCK_TENSORRT 1
#if TRTX_MOCK_TRT_MAJOR >= 8
#define TRTX_NX noexcept
#define TRTX_CE const
#else
#define TRTX_NX
#define TRTX_CE
#endif

struct cudnnContext;
struct cublasContext;

namespace nvinfer1 {

using AsciiChar = char;

enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4, kUINT8 = 5, kFP8 = 6 };

enum class TensorFormat : int32_t {
    kLINEAR = 0, kCHW2 = 1, kHWC8 = 2, kCHW4 = 3, kCHW16 = 4, kCHW32 = 5, kDHWC8 = 6, kCDHW32 = 7, kHWC = 8,
    kDLA_LINEAR = 9, kDLA_HWC4 = 10, kHWC16 = 11, kDHWC = 12
};
using PluginFormat = TensorFormat;

class Dims32 {
   public:
    static constexpr int32_t MAX_DIMS{8};
    int32_t nbDims;
    int32_t d[MAX_DIMS];
};
using Dims = Dims32;

class Dims2 : public Dims {
   public:
    Dims2() : Dims2(0, 0) {}
    Dims2(int32_t d0, int32_t d1) {
        nbDims = 2;
        d[0] = d0;
        d[1] = d1;
        for (int i = 2; i < MAX_DIMS; ++i) d[i] = 0;
    }
};
class DimsHW : public Dims2 {
   public:
    DimsHW() : Dims2() {}
    DimsHW(int32_t h, int32_t w) : Dims2(h, w) {}
};
class Dims3 : public Dims {
   public:
    Dims3() : Dims3(0, 0, 0) {}
    Dims3(int32_t d0, int32_t d1, int32_t d2) {
        nbDims = 3;
        d[0] = d0;
        d[1] = d1;
        d[2] = d2;
        for (int i = 3; i < MAX_DIMS; ++i) d[i] = 0;
    }
};
class Dims4 : public Dims {
   public:
    Dims4() : Dims4(0, 0, 0, 0) {}
    Dims4(int32_t d0, int32_t d1, int32_t d2, int32_t d3) {
        nbDims = 4;
        d[0] = d0;
        d[1] = d1;
        d[2] = d2;
        d[3] = d3;
        for (int i = 4; i < MAX_DIMS; ++i) d[i] = 0;
    }
};

struct PluginTensorDesc {
    Dims dims;
    DataType type;
    TensorFormat format;
    float scale;
};
struct DynamicPluginTensorDesc {
    PluginTensorDesc desc;
    Dims min;
    Dims max;
};

enum class PluginFieldType : int32_t {
    kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5, kCHAR = 6, kDIMS = 7, kUNKNOWN = 8
};
class PluginField {
   public:
    const AsciiChar* name;
    const void* data;
    PluginFieldType type;
    int32_t length;
    PluginField(const AsciiChar* const name_ = nullptr, const void* const data_ = nullptr,
                const PluginFieldType type_ = PluginFieldType::kUNKNOWN, int32_t length_ = 0) TRTX_NX
        : name(name_), data(data_), type(type_), length(length_) {}
};
struct PluginFieldCollection {
    int32_t nbFields;
    const PluginField* fields;
};

class IGpuAllocator;

// --- just enough of the builder API for the reference helper headers that are included by the files
//     under test (retinaface/common.hpp: loadWeights / addBatchNorm2d) to COMPILE; never instantiated ---
class Weights {
   public:
    DataType type;
    const void* values;
    int64_t count;
};
enum class ScaleMode : int32_t { kUNIFORM = 0, kCHANNEL = 1, kELEMENTWISE = 2 };
class ITensor;
class ILayer {
   public:
    virtual ITensor* getOutput(int32_t index) const TRTX_NX = 0;

   protected:
    virtual ~ILayer() TRTX_NX = default;
};
class IScaleLayer : public ILayer {};
class INetworkDefinition {
   public:
    virtual IScaleLayer* addScale(ITensor& input, ScaleMode mode, Weights shift, Weights scale, Weights power) TRTX_NX = 0;

   protected:
    virtual ~INetworkDefinition() TRTX_NX = default;
};

enum class DimensionOperation : int32_t { kSUM = 0, kPROD = 1, kMAX = 2, kMIN = 3, kSUB = 4, kEQUAL = 5, kLESS = 6, kFLOOR_DIV = 7, kCEIL_DIV = 8 };
class IDimensionExpr {
   public:
    virtual bool isConstant() const TRTX_NX = 0;
    virtual int32_t getConstantValue() const TRTX_NX = 0;

   protected:
    virtual ~IDimensionExpr() TRTX_NX = default;
};
class IExprBuilder {
   public:
    virtual const IDimensionExpr* constant(int32_t value) TRTX_NX = 0;
    virtual const IDimensionExpr* operation(DimensionOperation op, const IDimensionExpr& first,
                                            const IDimensionExpr& second) TRTX_NX = 0;

   protected:
    virtual ~IExprBuilder() TRTX_NX = default;
};
class DimsExprs {
   public:
    int32_t nbDims;
    const IDimensionExpr* d[Dims::MAX_DIMS];
};

// logger / profiler interfaces (the reference's yolov3-spp/Utils.h derives from them)
class ILogger {
   public:
    enum class Severity : int32_t { kINTERNAL_ERROR = 0, kERROR = 1, kWARNING = 2, kINFO = 3, kVERBOSE = 4 };
    virtual void log(Severity severity, const AsciiChar* msg) TRTX_NX = 0;
    virtual ~ILogger() = default;
};
class IProfiler {
   public:
    virtual void reportLayerTime(const char* layerName, float ms) TRTX_NX = 0;
    virtual ~IProfiler() = default;
};

// ---------------------------------------------------------------------------------------------
class IPluginV2 {
   public:
    virtual int32_t getTensorRTVersion() const TRTX_NX { return NV_TENSORRT_VERSION; }
    virtual const AsciiChar* getPluginType() const TRTX_NX = 0;
    virtual const AsciiChar* getPluginVersion() const TRTX_NX = 0;
    virtual int32_t getNbOutputs() const TRTX_NX = 0;
    virtual Dims getOutputDimensions(int32_t index, const Dims* inputs, int32_t nbInputDims) TRTX_NX = 0;
    virtual bool supportsFormat(DataType type, PluginFormat format) const TRTX_NX = 0;
    virtual void configureWithFormat(const Dims* inputDims, int32_t nbInputs, const Dims* outputDims, int32_t nbOutputs,
                                     DataType type, PluginFormat format, int32_t maxBatchSize) TRTX_NX = 0;
    virtual int32_t initialize() TRTX_NX = 0;
    virtual void terminate() TRTX_NX = 0;
    virtual size_t getWorkspaceSize(int32_t maxBatchSize) const TRTX_NX = 0;
    virtual int32_t enqueue(int32_t batchSize, const void* const* inputs, void* TRTX_CE* outputs, void* workspace,
                            cudaStream_t stream) TRTX_NX = 0;
    virtual size_t getSerializationSize() const TRTX_NX = 0;
    virtual void serialize(void* buffer) const TRTX_NX = 0;
    virtual void destroy() TRTX_NX = 0;
    virtual IPluginV2* clone() const TRTX_NX = 0;
    virtual void setPluginNamespace(const AsciiChar* pluginNamespace) TRTX_NX = 0;
    virtual const AsciiChar* getPluginNamespace() const TRTX_NX = 0;

    IPluginV2() = default;
    virtual ~IPluginV2() TRTX_NX = default;
};

class IPluginV2Ext : public IPluginV2 {
   public:
    virtual DataType getOutputDataType(int32_t index, const DataType* inputTypes, int32_t nbInputs) const TRTX_NX = 0;
    virtual bool isOutputBroadcastAcrossBatch(int32_t outputIndex, const bool* inputIsBroadcasted,
                                              int32_t nbInputs) const TRTX_NX = 0;
    virtual bool canBroadcastInputAcrossBatch(int32_t inputIndex) const TRTX_NX = 0;
    virtual void configurePlugin(const Dims* inputDims, int32_t nbInputs, const Dims* outputDims, int32_t nbOutputs,
                                 const DataType* inputTypes, const DataType* outputTypes, const bool* inputIsBroadcast,
                                 const bool* outputIsBroadcast, PluginFormat floatFormat, int32_t maxBatchSize) TRTX_NX = 0;
    virtual void attachToContext(cudnnContext*, cublasContext*, IGpuAllocator*) TRTX_NX {}
    virtual void detachFromContext() TRTX_NX {}
    IPluginV2Ext* clone() const TRTX_NX override = 0;

   protected:
    void configureWithFormat(const Dims*, int32_t, const Dims*, int32_t, DataType, PluginFormat, int32_t) TRTX_NX override {}
};

class IPluginV2IOExt : public IPluginV2Ext {
   public:
    virtual void configurePlugin(const PluginTensorDesc* in, int32_t nbInput, const PluginTensorDesc* out,
                                 int32_t nbOutput) TRTX_NX = 0;
    virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs,
                                           int32_t nbOutputs) const TRTX_NX = 0;

   private:
    void configurePlugin(const Dims*, int32_t, const Dims*, int32_t, const DataType*, const DataType*, const bool*,
                         const bool*, PluginFormat, int32_t) TRTX_NX final {}
    bool supportsFormat(DataType, PluginFormat) const TRTX_NX final { return false; }
};

class IPluginV2DynamicExt : public IPluginV2Ext {
   public:
    IPluginV2DynamicExt* clone() const TRTX_NX override = 0;
    virtual DimsExprs getOutputDimensions(int32_t outputIndex, const DimsExprs* inputs, int32_t nbInputs,
                                          IExprBuilder& exprBuilder) TRTX_NX = 0;
    virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs,
                                           int32_t nbOutputs) TRTX_NX = 0;
    virtual void configurePlugin(const DynamicPluginTensorDesc* in, int32_t nbInputs, const DynamicPluginTensorDesc* out,
                                 int32_t nbOutputs) TRTX_NX = 0;
    virtual size_t getWorkspaceSize(const PluginTensorDesc* inputs, int32_t nbInputs, const PluginTensorDesc* outputs,
                                    int32_t nbOutputs) const TRTX_NX = 0;
    virtual int32_t enqueue(const PluginTensorDesc* inputDesc, const PluginTensorDesc* outputDesc,
                            const void* const* inputs, void* const* outputs, void* workspace,
                            cudaStream_t stream) TRTX_NX = 0;  // `void* const*` in TensorRT 7 as well

   private:
    // implicit-batch entry points are sealed off, as in TensorRT
    Dims getOutputDimensions(int32_t, const Dims*, int32_t) TRTX_NX final { return Dims{-1, {}}; }
    bool isOutputBroadcastAcrossBatch(int32_t, const bool*, int32_t) const TRTX_NX final { return false; }
    bool canBroadcastInputAcrossBatch(int32_t) const TRTX_NX final { return true; }
    bool supportsFormat(DataType, PluginFormat) const TRTX_NX final { return false; }
    void configurePlugin(const Dims*, int32_t, const Dims*, int32_t, const DataType*, const DataType*, const bool*,
                         const bool*, PluginFormat, int32_t) TRTX_NX final {}
    size_t getWorkspaceSize(int32_t) const TRTX_NX final { return 0; }
    int32_t enqueue(int32_t, const void* const*, void* TRTX_CE*, void*, cudaStream_t) TRTX_NX final { return 1; }
};

class IPluginCreator {
   public:
    virtual int32_t getTensorRTVersion() const TRTX_NX { return NV_TENSORRT_VERSION; }
    virtual const AsciiChar* getPluginName() const TRTX_NX = 0;
    virtual const AsciiChar* getPluginVersion() const TRTX_NX = 0;
    virtual const PluginFieldCollection* getFieldNames() TRTX_NX = 0;
    virtual IPluginV2* createPlugin(const AsciiChar* name, const PluginFieldCollection* fc) TRTX_NX = 0;
    virtual IPluginV2* deserializePlugin(const AsciiChar* name, const void* serialData, size_t serialLength) TRTX_NX = 0;
    virtual void setPluginNamespace(const AsciiChar* pluginNamespace) TRTX_NX = 0;
    virtual const AsciiChar* getPluginNamespace() const TRTX_NX = 0;
    IPluginCreator() = default;
    virtual ~IPluginCreator() = default;
};

// minimal in-process registry (TensorRT's lives in libnvinfer)
class IPluginRegistry {
   public:
    bool registerCreator(IPluginCreator& creator, const AsciiChar* pluginNamespace) TRTX_NX {
        creators()[key(creator.getPluginName(), creator.getPluginVersion(), pluginNamespace)] = &creator;
        return true;
    }
    IPluginCreator* getPluginCreator(const AsciiChar* pluginName, const AsciiChar* pluginVersion,
                                     const AsciiChar* pluginNamespace = "") TRTX_NX {
        auto it = creators().find(key(pluginName, pluginVersion, pluginNamespace));
        return it == creators().end() ? nullptr : it->second;
    }

   private:
    static std::string key(const char* n, const char* v, const char* ns) {
        return std::string(ns ? ns : "") + "::" + (n ? n : "") + "::" + (v ? v : "");
    }
    static std::map<std::string, IPluginCreator*>& creators() {
        static std::map<std::string, IPluginCreator*> m;
        return m;
    }
};

template <typename T>
class PluginRegistrar {
   public:
    PluginRegistrar();

   private:
    T instance{};
};

}  // namespace nvinfer1

inline nvinfer1::IPluginRegistry* getPluginRegistry() TRTX_NX {
    static nvinfer1::IPluginRegistry r;
    return &r;
}

namespace nvinfer1 {
template <typename T>
PluginRegistrar<T>::PluginRegistrar() {
    getPluginRegistry()->registerCreator(instance, "");
}
}  // namespace nvinfer1

#ifdef TRTX_MOCK_NO_REGISTRAR  // host-only builds that include a plugin header without its .cu
#define REGISTER_TENSORRT_PLUGIN(name) static_assert(true, "")
#else
#define REGISTER_TENSORRT_PLUGIN(name) static nvinfer1::PluginRegistrar<name> pluginRegistrar##name {}
#endif

#endif  // TRTX_MOCK_NVINFER_H
