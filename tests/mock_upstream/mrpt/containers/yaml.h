#pragma once
// the few operations of mrpt::containers::yaml the adapters use
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace mrpt::containers
{
class yaml
{
   public:
    using sequence_t = std::vector<yaml>;
    using map_t      = std::map<std::string, yaml>;
    yaml() = default;
    yaml(const char* s) : scalar_(s), kind_(Scalar) {}
    yaml(const std::string& s) : scalar_(s), kind_(Scalar) {}
    yaml(double v) : kind_(Scalar) { std::ostringstream o; o.precision(17); o << v; scalar_ = o.str(); }
    yaml(int v) : scalar_(std::to_string(v)), kind_(Scalar) {}
    yaml(bool v) : scalar_(v ? "true" : "false"), kind_(Scalar) {}
    static yaml Map() { yaml y; y.kind_ = MapK; return y; }
    static yaml Sequence() { yaml y; y.kind_ = Seq; return y; }
    bool isMap() const { return kind_ == MapK; }
    bool isSequence() const { return kind_ == Seq; }
    bool isScalar() const { return kind_ == Scalar; }
    bool has(const std::string& k) const { return kind_ == MapK && map_.count(k) != 0; }
    yaml& operator[](const std::string& k) { kind_ = MapK; return map_[k]; }
    const yaml& operator[](const std::string& k) const
    {
        auto it = map_.find(k);
        if (kind_ != MapK || it == map_.end()) throw std::out_of_range("yaml: no key " + k);
        return it->second;
    }
    void push_back(const yaml& v) { kind_ = Seq; seq_.push_back(v); }
    const sequence_t& asSequence() const { if (kind_ != Seq) throw std::logic_error("yaml: not a sequence"); return seq_; }
    const map_t&      asMap() const { if (kind_ != MapK) throw std::logic_error("yaml: not a map"); return map_; }
    template <typename T> T as() const
    {
        if (kind_ != Scalar) throw std::logic_error("yaml: not a scalar");
        if constexpr (std::is_same_v<T, std::string>) return scalar_;
        else if constexpr (std::is_same_v<T, bool>) return scalar_ == "true" || scalar_ == "True" || scalar_ == "1";
        else { std::istringstream i(scalar_); T v{}; i >> v; if (i.fail()) throw std::logic_error("yaml: bad number " + scalar_); return v; }
    }
    template <typename T> T getOrDefault(const std::string& k, const T& d) const { return has(k) ? (*this)[k].as<T>() : d; }
   private:
    enum Kind { Null, Scalar, Seq, MapK };
    std::string scalar_;
    sequence_t  seq_;
    map_t       map_;
    Kind        kind_ = Null;
};
}  // namespace mrpt::containers
