// yaml_lite.cpp -- see yaml_lite.h
#include "yaml_lite.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace yaml_lite
{
namespace
{
struct Line
{
    int         indent;
    std::string text;
    int         number;
};

std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) a++;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) b--;
    return s.substr(a, b - a);
}

// removes a trailing '# comment' that is outside quotes
std::string strip_comment(const std::string& s)
{
    char quote = 0;
    for (size_t i = 0; i < s.size(); i++)
    {
        const char c = s[i];
        if (quote)
        {
            if (c == quote) quote = 0;
        }
        else if (c == '"' || c == '\'')
            quote = c;
        else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t'))
            return s.substr(0, i);
    }
    return s;
}

std::string run_command(const std::string& cmd, const Options& opt)
{
    std::istringstream is(cmd);
    std::string        tool;
    is >> tool;
    if (tool == "mola-dir")
    {
        std::string name;
        is >> name;
        auto it = opt.module_dirs.find(name);
        if (it != opt.module_dirs.end()) return it->second;
        std::string env = "MOLA_DIR_" + name;
        for (auto& ch : env)
            if (ch == '-') ch = '_';
        if (const char* e = std::getenv(env.c_str())) return e;
        throw std::runtime_error("yaml: cannot resolve $(mola-dir " + name + "): no directory registered");
    }
    if (!opt.allow_shell) throw std::runtime_error("yaml: $(" + cmd + ") not allowed");
    std::string out;
    FILE*       f = popen(cmd.c_str(), "r");
    if (!f) throw std::runtime_error("yaml: cannot run $(" + cmd + ")");
    char buf[256];
    while (fgets(buf, sizeof(buf), f)) out += buf;
    pclose(f);
    return trim(out);
}

// expands $(cmd) and ${VAR}; leaves $include{...} for the value parser
std::string expand(const std::string& s, const Options& opt)
{
    std::string o;
    for (size_t i = 0; i < s.size();)
    {
        if (s[i] == '$' && i + 1 < s.size() && (s[i + 1] == '(' || s[i + 1] == '{'))
        {
            const char   close = s[i + 1] == '(' ? ')' : '}';
            const size_t e = s.find(close, i + 2);
            if (e == std::string::npos) throw std::runtime_error("yaml: unterminated $ expression in: " + s);
            const std::string inner = s.substr(i + 2, e - i - 2);
            if (close == ')')
                o += run_command(expand(inner, opt), opt);
            else
            {
                const char* v = std::getenv(inner.c_str());
                if (!v) throw std::runtime_error("yaml: environment variable ${" + inner + "} not set");
                o += v;
            }
            i = e + 1;
        }
        else
            o += s[i++];
    }
    return o;
}

std::string unquote(const std::string& s)
{
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
        return s.substr(1, s.size() - 2);
    return s;
}

struct Parser
{
    std::vector<Line> lines;
    size_t            pos = 0;
    const Options&    opt;
    explicit Parser(const Options& o) : opt(o) {}

    Node value_from_text(const std::string& raw)
    {
        const std::string v = trim(raw);
        // $include{path}: the included file's tree becomes the value
        const std::string inc = "$include{";
        if (v.compare(0, inc.size(), inc) == 0)
        {
            // the path itself may hold $(...) and ${...}: find the matching close brace
            int    depth = 0;
            size_t e = std::string::npos;
            for (size_t i = inc.size() - 1; i < v.size(); i++)
            {
                if (v[i] == '{' || v[i] == '(') depth++;
                if (v[i] == '}' || v[i] == ')')
                {
                    depth--;
                    if (depth == 0)
                    {
                        e = i;
                        break;
                    }
                }
            }
            if (e == std::string::npos) throw std::runtime_error("yaml: unterminated $include{ in: " + v);
            std::string path = expand(v.substr(inc.size(), e - inc.size()), opt);
            if (!path.empty() && path[0] != '/' && !opt.base_dir.empty()) path = opt.base_dir + "/" + path;
            return parse_file(path, opt);
        }
        Node n;
        if (v.empty() || v == "~" || v == "null")
            return n;
        if (v == "[]")
        {
            n.type = Node::Seq;
            return n;
        }
        if (v == "{}")
        {
            n.type = Node::Map;
            return n;
        }
        n.type = Node::Scalar;
        n.scalar = unquote(expand(v, opt));
        return n;
    }

    static bool is_seq_item(const std::string& t) { return t == "-" || t.compare(0, 2, "- ") == 0; }

    // position of the ':' that ends a map key, or npos
    static size_t key_colon(const std::string& t)
    {
        char quote = 0;
        int  depth = 0;
        for (size_t i = 0; i < t.size(); i++)
        {
            const char c = t[i];
            if (quote)
            {
                if (c == quote) quote = 0;
                continue;
            }
            if (c == '"' || c == '\'') quote = c;
            if (c == '{' || c == '(') depth++;
            if (c == '}' || c == ')') depth--;
            if (c == ':' && depth == 0 && (i + 1 == t.size() || t[i + 1] == ' ' || t[i + 1] == '\t')) return i;
        }
        return std::string::npos;
    }

    Node parse_block(int indent)
    {
        Node n;
        if (pos >= lines.size()) return n;
        if (is_seq_item(lines[pos].text))
        {
            n.type = Node::Seq;
            while (pos < lines.size() && lines[pos].indent == indent && is_seq_item(lines[pos].text))
            {
                std::string rest = lines[pos].text.size() > 1 ? lines[pos].text.substr(2) : "";
                const size_t lead = rest.find_first_not_of(' ');
                rest = trim(rest);
                if (rest.empty())
                {
                    pos++;
                    if (pos < lines.size() && lines[pos].indent > indent)
                        n.seq.push_back(parse_block(lines[pos].indent));
                    else
                        n.seq.push_back(Node());
                }
                else if (key_colon(rest) != std::string::npos)
                {
                    // "- key: value": a map whose first key sits after the dash
                    const int child = indent + 2 + (int)(lead == std::string::npos ? 0 : lead);
                    lines[pos].indent = child;
                    lines[pos].text = rest;
                    n.seq.push_back(parse_block(child));
                }
                else
                {
                    n.seq.push_back(value_from_text(rest));
                    pos++;
                }
            }
            return n;
        }
        n.type = Node::Map;
        while (pos < lines.size() && lines[pos].indent == indent && !is_seq_item(lines[pos].text))
        {
            const std::string& t = lines[pos].text;
            const size_t       c = key_colon(t);
            if (c == std::string::npos && t.compare(0, 9, "$include{") == 0)
            {
                // a whole-line $include{} inside a map: merge the file's entries
                Node inc = value_from_text(t);
                if (!inc.isMap() && !inc.isNull())
                    throw std::runtime_error("yaml: line " + std::to_string(lines[pos].number) +
                                             ": $include{} inside a map must yield a map");
                for (auto& kv : inc.map) n.map.push_back(std::move(kv));
                pos++;
                continue;
            }
            if (c == std::string::npos)
                throw std::runtime_error("yaml: line " + std::to_string(lines[pos].number) +
                                         ": expected 'key: value', got: " + t);
            const std::string key = unquote(trim(t.substr(0, c)));
            const std::string val = trim(t.substr(c + 1));
            pos++;
            Node child;
            if (val.empty())
            {
                if (pos < lines.size() &&
                    (lines[pos].indent > indent ||
                     (lines[pos].indent == indent && is_seq_item(lines[pos].text))))
                    child = parse_block(lines[pos].indent);
            }
            else
                child = value_from_text(val);
            n.map.emplace_back(key, std::move(child));
        }
        if (pos < lines.size() && lines[pos].indent > indent)
            throw std::runtime_error("yaml: line " + std::to_string(lines[pos].number) + ": bad indentation");
        return n;
    }
};
}  // namespace

bool Node::has(const std::string& key) const
{
    if (type != Map) return false;
    for (const auto& kv : map)
        if (kv.first == key) return true;
    return false;
}

const Node& Node::operator[](const std::string& key) const
{
    static const Node null_node;
    if (type == Map)
        for (const auto& kv : map)
            if (kv.first == key) return kv.second;
    return null_node;
}

const Node& Node::at(const std::string& key) const
{
    if (!has(key)) throw std::runtime_error("yaml: required entry `" + key + "` not found");
    return (*this)[key];
}

std::string Node::as_string() const
{
    if (type != Scalar) throw std::runtime_error("yaml: expected a scalar");
    return scalar;
}
double Node::as_double() const
{
    const std::string s = as_string();
    char*             end = nullptr;
    const double      v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end != 0) throw std::runtime_error("yaml: `" + s + "` is not a number");
    return v;
}
long Node::as_int() const
{
    const std::string s = as_string();
    char*             end = nullptr;
    const long        v = std::strtol(s.c_str(), &end, 10);
    if (end != s.c_str() && *end == 0) return v;
    const double d = as_double();
    return (long)d;
}
bool Node::as_bool() const
{
    const std::string s = as_string();
    if (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on" || s == "1") return true;
    if (s == "false" || s == "False" || s == "FALSE" || s == "no" || s == "off" || s == "0") return false;
    throw std::runtime_error("yaml: `" + s + "` is not a boolean");
}

std::string Node::dump(int indent) const
{
    std::string       o;
    const std::string pad(indent, ' ');
    switch (type)
    {
        case Null: return "~";
        case Scalar: return scalar;
        case Map:
            for (const auto& kv : map)
            {
                o += pad + kv.first + ":";
                if (kv.second.type == Scalar || kv.second.type == Null)
                    o += " " + kv.second.dump() + "\n";
                else
                    o += "\n" + kv.second.dump(indent + 2);
            }
            return o;
        case Seq:
            for (const auto& it : seq)
            {
                if (it.type == Scalar || it.type == Null)
                    o += pad + "- " + it.dump() + "\n";
                else
                    o += pad + "-\n" + it.dump(indent + 2);
            }
            return o;
    }
    return o;
}

Node parse(const std::string& text, const Options& opt)
{
    Parser             p(opt);
    std::istringstream is(text);
    std::string        raw;
    int                num = 0;
    while (std::getline(is, raw))
    {
        num++;
        std::string s = strip_comment(raw);
        if (trim(s).empty()) continue;
        if (trim(s) == "---") continue;
        int indent = 0;
        while (indent < (int)s.size() && s[indent] == ' ') indent++;
        if (indent < (int)s.size() && s[indent] == '\t')
            throw std::runtime_error("yaml: line " + std::to_string(num) + ": tab indentation");
        p.lines.push_back({indent, trim(s), num});
    }
    if (p.lines.empty()) return Node();
    Node n = p.parse_block(p.lines[0].indent);
    if (p.pos != p.lines.size())
        throw std::runtime_error("yaml: line " + std::to_string(p.lines[p.pos].number) +
                                 ": unexpected content");
    return n;
}

Node parse_file(const std::string& path, Options opt)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("yaml: cannot open file `" + path + "`");
    std::stringstream ss;
    ss << f.rdbuf();
    const size_t slash = path.find_last_of('/');
    opt.base_dir = (slash == std::string::npos) ? "." : path.substr(0, slash);
    return parse(ss.str(), opt);
}

}  // namespace yaml_lite
