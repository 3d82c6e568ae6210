#pragma once
// the shape of MRPT's run-time class registry (mrpt::rtti), reduced to a name -> factory map
#include <functional>
#include <map>
#include <memory>
#include <string>
namespace mrpt::rtti
{
class CObject
{
   public:
    using Ptr = std::shared_ptr<CObject>;
    virtual ~CObject() = default;
    virtual const char* className() const { return "CObject"; }
};
struct TRuntimeClassId { const char* className; std::function<CObject::Ptr()> create; };
inline std::map<std::string, const TRuntimeClassId*>& registry() { static std::map<std::string, const TRuntimeClassId*> r; return r; }
inline void registerClass(const TRuntimeClassId* id) { registry()[id->className] = id; }
inline CObject::Ptr classFactory(const std::string& name)
{
    auto it = registry().find(name);
    return it == registry().end() ? nullptr : it->second->create();
}
}  // namespace mrpt::rtti
namespace mrpt
{
template <typename T> struct ptr_cast
{
    template <typename U> static std::shared_ptr<T> from(const std::shared_ptr<U>& p) { return std::dynamic_pointer_cast<T>(p); }
};
}  // namespace mrpt
#define DEFINE_MRPT_OBJECT(cls, NS)                                                     \
   public:                                                                              \
    using Ptr = std::shared_ptr<cls>;                                                   \
    static const mrpt::rtti::TRuntimeClassId& GetRuntimeClassIdStatic();                \
    static std::shared_ptr<cls> Create() { return std::make_shared<cls>(); }            \
    const char* className() const override { return #NS "::" #cls; }
#define IMPLEMENTS_MRPT_OBJECT(cls, base, NS)                                           \
    const mrpt::rtti::TRuntimeClassId& NS::cls::GetRuntimeClassIdStatic()               \
    {                                                                                   \
        static const mrpt::rtti::TRuntimeClassId id{#NS "::" #cls, [] { return std::static_pointer_cast<mrpt::rtti::CObject>(std::make_shared<NS::cls>()); }}; \
        return id;                                                                      \
    }
#define CLASS_ID(T) (&T::GetRuntimeClassIdStatic())
