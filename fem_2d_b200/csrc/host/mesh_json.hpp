// Minimal JSON reader for the reference's mesh-file format (mesh.rs:100-141, parse_* at mesh.rs:1683-1783):
//   { "Elements": [ {"materials": [eps_re, eps_im, mu_re, mu_im], "node_ids": [n0,n1,n2,n3]}, ... ], "Nodes": [[x,y], ...] }
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>

#include "mesh.hpp"

namespace fem2d {
namespace json {
struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    double num = 0; bool b = false; std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;
    const Value* get(const std::string& k) const { for (auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
};
class Parser {
    const std::string& s; size_t p = 0;
    [[noreturn]] void fail(const char* m) const { throw MeshError(MeshError::BadMeshFile, p, std::string("Unable to parse Mesh File as JSON: ") + m); }
    void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) p++; }
  public:
    explicit Parser(const std::string& src) : s(src) {}
    Value parse() { Value v = value(); ws(); if (p != s.size()) fail("trailing characters"); return v; }
    Value value() {
        ws(); if (p >= s.size()) fail("unexpected end");
        Value v; char c = s[p];
        if (c == '{') {
            v.type = Value::Object; p++; ws();
            if (s[p] == '}') { p++; return v; }
            for (;;) {
                ws(); if (s[p] != '"') fail("expected key");
                std::string k = string(); ws(); if (s[p++] != ':') fail("expected ':'");
                v.obj.emplace_back(k, value()); ws();
                if (s[p] == ',') { p++; continue; }
                if (s[p] == '}') { p++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.type = Value::Array; p++; ws();
            if (s[p] == ']') { p++; return v; }
            for (;;) {
                v.arr.push_back(value()); ws();
                if (s[p] == ',') { p++; continue; }
                if (s[p] == ']') { p++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') { v.type = Value::String; v.str = string(); }
        else if (!s.compare(p, 4, "true")) { v.type = Value::Bool; v.b = true; p += 4; }
        else if (!s.compare(p, 5, "false")) { v.type = Value::Bool; p += 5; }
        else if (!s.compare(p, 4, "null")) { p += 4; }
        else {
            char* end = nullptr; v.num = std::strtod(s.c_str() + p, &end);
            if (end == s.c_str() + p) fail("bad number");
            v.type = Value::Number; p = end - s.c_str();
        }
        return v;
    }
    std::string string() {
        std::string out; p++;
        while (p < s.size() && s[p] != '"') { if (s[p] == '\\' && p + 1 < s.size()) p++; out.push_back(s[p++]); }
        if (p >= s.size()) fail("unterminated string");
        p++; return out;
    }
};
}  // namespace json

inline Mesh Mesh::from_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw MeshError(MeshError::BadMeshFile, 0, "cannot open mesh file: " + path);
    std::stringstream ss; ss << f.rdbuf();
    const std::string text = ss.str();
    json::Value root = json::Parser(text).parse();
    const json::Value* els = root.get("Elements"); const json::Value* nds = root.get("Nodes");
    if (!els || els->type != json::Value::Array) throw MeshError(MeshError::BadMeshFile, 0, "Elements must be an Array!");
    if (!nds || nds->type != json::Value::Array) throw MeshError(MeshError::BadMeshFile, 0, "Nodes must be an Array!");
    std::vector<double> mats, xy; std::vector<int64_t> nids;
    for (auto& e : els->arr) {
        const json::Value* m = e.get("materials"); const json::Value* n = e.get("node_ids");
        if (!n || n->type != json::Value::Array || n->arr.size() != 4) throw MeshError(MeshError::BadMeshFile, 0, "Elements Array of node_ids must have a length of 4!");
        if (!m || m->type != json::Value::Array || m->arr.size() != 4) throw MeshError(MeshError::BadMeshFile, 0, "Elements Array of materials must have a length of 4!");
        for (auto& v : m->arr) { if (v.type != json::Value::Number) throw MeshError(MeshError::BadMeshFile, 0, "Element materials must be numerical values"); mats.push_back(v.num); }
        for (auto& v : n->arr) {
            if (v.type != json::Value::Number || v.num < 0 || v.num != std::floor(v.num)) throw MeshError(MeshError::BadMeshFile, 0, "node_ids must be positive integers!");
            nids.push_back((int64_t)v.num);
        }
    }
    for (auto& n : nds->arr) {
        if (n.type != json::Value::Array || n.arr.size() != 2) throw MeshError(MeshError::BadMeshFile, 0, "nodes must be arrays of length 2!");
        for (auto& v : n.arr) { if (v.type != json::Value::Number) throw MeshError(MeshError::BadMeshFile, 0, "nodes must be composed of numerical values!"); xy.push_back(v.num); }
    }
    return from_arrays(els->arr.size(), mats.data(), nids.data(), nds->arr.size(), xy.data());
}
}  // namespace fem2d
