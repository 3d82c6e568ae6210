// yaml_lite.h -- the subset of YAML + the mola-yaml text extensions that the
// reference's parameter files use (params/*.yaml; consumer macros
// YAML_LOAD_REQ / YAML_LOAD_OPT / YAML_LOAD_OPT_DEG / ENSURE_YAML_ENTRY_EXISTS,
// LidarOdometry.cpp:20,63,77-86,105-128):
//   block maps and block sequences by indentation, plain / quoted scalars,
//   '#' comments, and the preprocessor forms  $include{path}  $(command)
//   ${ENV_VAR}  (kitti-default.yaml:43,46,50).
// `$(mola-dir NAME)` is resolved from a table of module directories instead of
// running the (absent) mola-dir tool; other commands go through popen().
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace yaml_lite
{
struct Node
{
    enum Type { Null, Scalar, Map, Seq } type = Null;
    std::string                               scalar;
    std::vector<std::pair<std::string, Node>> map;
    std::vector<Node>                         seq;

    bool        has(const std::string& key) const;
    const Node& at(const std::string& key) const;  // throws like ENSURE_YAML_ENTRY_EXISTS
    const Node& operator[](const std::string& key) const;  // Null node if absent
    bool        isNull() const { return type == Null; }
    bool        isMap() const { return type == Map; }
    bool        isSeq() const { return type == Seq; }
    bool        isScalar() const { return type == Scalar; }

    std::string as_string() const;
    double      as_double() const;
    long        as_int() const;
    bool        as_bool() const;

    // YAML_LOAD_OPT semantics: leave `v` untouched when the key is absent
    template <typename T>
    void load_opt(const std::string& key, T& v) const;
    // YAML_LOAD_REQ semantics: throw when absent
    template <typename T>
    void load_req(const std::string& key, T& v) const;

    std::string dump(int indent = 0) const;
};

struct Options
{
    // `$(mola-dir NAME)` -> directory
    std::map<std::string, std::string> module_dirs;
    // base directory for relative $include{} paths
    std::string base_dir;
    bool        allow_shell = true;
};

Node parse(const std::string& text, const Options& opt = Options());
Node parse_file(const std::string& path, Options opt = Options());

template <typename T>
struct Conv;
template <>
struct Conv<double> { static double get(const Node& n) { return n.as_double(); } };
template <>
struct Conv<float> { static float get(const Node& n) { return (float)n.as_double(); } };
template <>
struct Conv<int> { static int get(const Node& n) { return (int)n.as_int(); } };
template <>
struct Conv<unsigned int> { static unsigned int get(const Node& n) { return (unsigned int)n.as_int(); } };
template <>
struct Conv<bool> { static bool get(const Node& n) { return n.as_bool(); } };
template <>
struct Conv<std::string> { static std::string get(const Node& n) { return n.as_string(); } };

template <typename T>
void Node::load_opt(const std::string& key, T& v) const
{
    if (has(key)) v = Conv<T>::get((*this)[key]);
}
template <typename T>
void Node::load_req(const std::string& key, T& v) const
{
    v = Conv<T>::get(at(key));
}

}  // namespace yaml_lite
