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

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <utility>

#define NV_TENSORRT_MAJOR 8
#define NV_TENSORRT_MINOR 6
#define NV_TENSORRT_PATCH 1
#define NV_TENSORRT_VERSION 8601
#define TRTX_MOCK_TENSORRT 1

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
                const PluginFieldType type_ = PluginFieldType::kUNKNOWN, int32_t length_ = 0) noexcept
        : name(name_), data(data_), type(type_), length(length_) {}
};
struct PluginFieldCollection {
    int32_t nbFields;
    const PluginField* fields;
};

class IGpuAllocator;

enum class DimensionOperation : int32_t { kSUM = 0, kPROD = 1, kMAX = 2, kMIN = 3, kSUB = 4, kEQUAL = 5, kLESS = 6, kFLOOR_DIV = 7, kCEIL_DIV = 8 };
class IDimensionExpr {
   public:
    virtual bool isConstant() const noexcept = 0;
    virtual int32_t getConstantValue() const noexcept = 0;

   protected:
    virtual ~IDimensionExpr() noexcept = default;
};
class IExprBuilder {
   public:
    virtual const IDimensionExpr* constant(int32_t value) noexcept = 0;
    virtual const IDimensionExpr* operation(DimensionOperation op, const IDimensionExpr& first,
                                            const IDimensionExpr& second) noexcept = 0;

   protected:
    virtual ~IExprBuilder() noexcept = default;
};
class DimsExprs {
   public:
    int32_t nbDims;
    const IDimensionExpr* d[Dims::MAX_DIMS];
};

// ---------------------------------------------------------------------------------------------
class IPluginV2 {
   public:
    virtual int32_t getTensorRTVersion() const noexcept { return NV_TENSORRT_VERSION; }
    virtual const AsciiChar* getPluginType() const noexcept = 0;
    virtual const AsciiChar* getPluginVersion() const noexcept = 0;
    virtual int32_t getNbOutputs() const noexcept = 0;
    virtual Dims getOutputDimensions(int32_t index, const Dims* inputs, int32_t nbInputDims) noexcept = 0;
    virtual bool supportsFormat(DataType type, PluginFormat format) const noexcept = 0;
    virtual void configureWithFormat(const Dims* inputDims, int32_t nbInputs, const Dims* outputDims, int32_t nbOutputs,
                                     DataType type, PluginFormat format, int32_t maxBatchSize) noexcept = 0;
    virtual int32_t initialize() noexcept = 0;
    virtual void terminate() noexcept = 0;
    virtual size_t getWorkspaceSize(int32_t maxBatchSize) const noexcept = 0;
    virtual int32_t enqueue(int32_t batchSize, const void* const* inputs, void* const* outputs, void* workspace,
                            cudaStream_t stream) noexcept = 0;
    virtual size_t getSerializationSize() const noexcept = 0;
    virtual void serialize(void* buffer) const noexcept = 0;
    virtual void destroy() noexcept = 0;
    virtual IPluginV2* clone() const noexcept = 0;
    virtual void setPluginNamespace(const AsciiChar* pluginNamespace) noexcept = 0;
    virtual const AsciiChar* getPluginNamespace() const noexcept = 0;

    IPluginV2() = default;
    virtual ~IPluginV2() noexcept = default;
};

class IPluginV2Ext : public IPluginV2 {
   public:
    virtual DataType getOutputDataType(int32_t index, const DataType* inputTypes, int32_t nbInputs) const noexcept = 0;
    virtual bool isOutputBroadcastAcrossBatch(int32_t outputIndex, const bool* inputIsBroadcasted,
                                              int32_t nbInputs) const noexcept = 0;
    virtual bool canBroadcastInputAcrossBatch(int32_t inputIndex) const noexcept = 0;
    virtual void configurePlugin(const Dims* inputDims, int32_t nbInputs, const Dims* outputDims, int32_t nbOutputs,
                                 const DataType* inputTypes, const DataType* outputTypes, const bool* inputIsBroadcast,
                                 const bool* outputIsBroadcast, PluginFormat floatFormat, int32_t maxBatchSize) noexcept = 0;
    virtual void attachToContext(cudnnContext*, cublasContext*, IGpuAllocator*) noexcept {}
    virtual void detachFromContext() noexcept {}
    IPluginV2Ext* clone() const noexcept override = 0;

   protected:
    void configureWithFormat(const Dims*, int32_t, const Dims*, int32_t, DataType, PluginFormat, int32_t) noexcept override {}
};

class IPluginV2IOExt : public IPluginV2Ext {
   public:
    virtual void configurePlugin(const PluginTensorDesc* in, int32_t nbInput, const PluginTensorDesc* out,
                                 int32_t nbOutput) noexcept = 0;
    virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs,
                                           int32_t nbOutputs) const noexcept = 0;

   private:
    void configurePlugin(const Dims*, int32_t, const Dims*, int32_t, const DataType*, const DataType*, const bool*,
                         const bool*, PluginFormat, int32_t) noexcept final {}
    bool supportsFormat(DataType, PluginFormat) const noexcept final { return false; }
};

class IPluginV2DynamicExt : public IPluginV2Ext {
   public:
    IPluginV2DynamicExt* clone() const noexcept override = 0;
    virtual DimsExprs getOutputDimensions(int32_t outputIndex, const DimsExprs* inputs, int32_t nbInputs,
                                          IExprBuilder& exprBuilder) noexcept = 0;
    virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc* inOut, int32_t nbInputs,
                                           int32_t nbOutputs) noexcept = 0;
    virtual void configurePlugin(const DynamicPluginTensorDesc* in, int32_t nbInputs, const DynamicPluginTensorDesc* out,
                                 int32_t nbOutputs) noexcept = 0;
    virtual size_t getWorkspaceSize(const PluginTensorDesc* inputs, int32_t nbInputs, const PluginTensorDesc* outputs,
                                    int32_t nbOutputs) const noexcept = 0;
    virtual int32_t enqueue(const PluginTensorDesc* inputDesc, const PluginTensorDesc* outputDesc,
                            const void* const* inputs, void* const* outputs, void* workspace,
                            cudaStream_t stream) noexcept = 0;

   private:
    // implicit-batch entry points are sealed off, as in TensorRT
    Dims getOutputDimensions(int32_t, const Dims*, int32_t) noexcept final { return Dims{-1, {}}; }
    bool isOutputBroadcastAcrossBatch(int32_t, const bool*, int32_t) const noexcept final { return false; }
    bool canBroadcastInputAcrossBatch(int32_t) const noexcept final { return true; }
    bool supportsFormat(DataType, PluginFormat) const noexcept final { return false; }
    void configurePlugin(const Dims*, int32_t, const Dims*, int32_t, const DataType*, const DataType*, const bool*,
                         const bool*, PluginFormat, int32_t) noexcept final {}
    size_t getWorkspaceSize(int32_t) const noexcept final { return 0; }
    int32_t enqueue(int32_t, const void* const*, void* const*, void*, cudaStream_t) noexcept final { return 1; }
};

class IPluginCreator {
   public:
    virtual int32_t getTensorRTVersion() const noexcept { return NV_TENSORRT_VERSION; }
    virtual const AsciiChar* getPluginName() const noexcept = 0;
    virtual const AsciiChar* getPluginVersion() const noexcept = 0;
    virtual const PluginFieldCollection* getFieldNames() noexcept = 0;
    virtual IPluginV2* createPlugin(const AsciiChar* name, const PluginFieldCollection* fc) noexcept = 0;
    virtual IPluginV2* deserializePlugin(const AsciiChar* name, const void* serialData, size_t serialLength) noexcept = 0;
    virtual void setPluginNamespace(const AsciiChar* pluginNamespace) noexcept = 0;
    virtual const AsciiChar* getPluginNamespace() const noexcept = 0;
    IPluginCreator() = default;
    virtual ~IPluginCreator() = default;
};

// minimal in-process registry (TensorRT's lives in libnvinfer)
class IPluginRegistry {
   public:
    bool registerCreator(IPluginCreator& creator, const AsciiChar* pluginNamespace) noexcept {
        creators()[key(creator.getPluginName(), creator.getPluginVersion(), pluginNamespace)] = &creator;
        return true;
    }
    IPluginCreator* getPluginCreator(const AsciiChar* pluginName, const AsciiChar* pluginVersion,
                                     const AsciiChar* pluginNamespace = "") noexcept {
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

inline nvinfer1::IPluginRegistry* getPluginRegistry() noexcept {
    static nvinfer1::IPluginRegistry r;
    return &r;
}

namespace nvinfer1 {
template <typename T>
PluginRegistrar<T>::PluginRegistrar() {
    getPluginRegistry()->registerCreator(instance, "");
}
}  // namespace nvinfer1

#define REGISTER_TENSORRT_PLUGIN(name) static nvinfer1::PluginRegistrar<name> pluginRegistrar##name {}

#endif  // TRTX_MOCK_NVINFER_H
