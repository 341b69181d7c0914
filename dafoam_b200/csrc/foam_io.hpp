// Host-side readers: OpenFOAM dictionaries / polyMesh (ASCII and `format binary`) and a minimal JSON
// parser for the DAOPTION dict.  Replaces, for the adjoint hot path, what the reference gets from
// OpenFOAM's IOobject/dictionary/polyMesh readers (reference src/adjoint/DASolver/DASolver.C:58-64,
// src/include/createMeshPython.H) and from DAUtility::pyDict2OFDict (reference
// src/adjoint/DAUtility/DAUtility.C:24-280).
#pragma once
#include <thread>
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dab
{

struct Error : std::runtime_error
{
    explicit Error(const std::string& m) : std::runtime_error(m) {}
};

namespace detail
{
// host threads for the set-up loops whose iterations are independent (DAB_HOST_THREADS overrides; default: the hardware
// concurrency, at most 64).  fn(threadIndex, begin, end); exceptions are re-thrown on the calling thread.
inline int hostThreads()
{
    if (const char* e = getenv("DAB_HOST_THREADS")) return std::max(1, atoi(e));
    const unsigned h = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(h ? h : 1u, 64u));
}
template <class Fn>
void parallelFor(int n, int nThreads, Fn fn)
{
    nThreads = std::max(1, std::min(nThreads, n / 4096 + 1));
    if (nThreads == 1)
    {
        fn(0, 0, n);
        return;
    }
    std::vector<std::thread> th;
    std::vector<std::string> err(nThreads);
    for (int t = 0; t < nThreads; t++)
        th.emplace_back([&, t]() {
            const int b = (int)((int64_t)n * t / nThreads), e = (int)((int64_t)n * (t + 1) / nThreads);
            try { fn(t, b, e); }
            catch (const std::exception& ex) { err[t] = ex.what(); if (err[t].empty()) err[t] = "error"; }
        });
    for (auto& x : th) x.join();
    for (auto& m : err)
        if (!m.empty()) throw Error(m);
}

} // namespace detail

inline std::string readFile(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open " + path);
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::string s((size_t)n, '\0');
    if (n > 0 && fread(&s[0], 1, (size_t)n, f) != (size_t)n)
    {
        fclose(f);
        throw Error("short read " + path);
    }
    fclose(f);
    return s;
}

inline bool fileExists(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

// ------------------------------------------------------------------------------------------------
// OpenFOAM dictionary
// ------------------------------------------------------------------------------------------------
struct Dict
{
    std::map<std::string, std::vector<std::string>> entries; // keyword -> tokens up to ';'
    std::map<std::string, Dict> subs;
    std::vector<std::string> order; // sub-dictionary names in file order

    bool has(const std::string& k) const { return entries.count(k) != 0; }
    bool hasSub(const std::string& k) const { return subs.count(k) != 0; }
    const Dict& sub(const std::string& k) const
    {
        auto it = subs.find(k);
        if (it == subs.end()) throw Error("missing sub-dictionary " + k);
        return it->second;
    }
    const std::vector<std::string>& tokens(const std::string& k) const
    {
        auto it = entries.find(k);
        if (it == entries.end()) throw Error("missing keyword " + k);
        return it->second;
    }
    std::string word(const std::string& k) const { return tokens(k).at(0); }
    std::string wordOr(const std::string& k, const std::string& d) const { return has(k) ? word(k) : d; }
    double scalar(const std::string& k) const
    {
        // `nu 1.5e-5;` or `nu [0 2 -1 0 0 0 0] 1.5e-5;` or `nu nu [dims] value;`
        const auto& t = tokens(k);
        return atof(t.back().c_str());
    }
    double scalarOr(const std::string& k, double d) const { return has(k) ? scalar(k) : d; }
    std::string joined(const std::string& k) const
    {
        std::string s;
        for (const auto& t : tokens(k)) s += (s.empty() ? "" : " ") + t;
        return s;
    }
    // `uniform 1.0`, `uniform (a b c)`: returns number of components parsed
    int uniform(const std::string& k, double* out) const
    {
        const auto& t = tokens(k);
        size_t i = 0;
        if (t.at(0) == "uniform") i = 1;
        int n = 0;
        for (; i < t.size(); i++)
        {
            if (t[i] == "(" || t[i] == ")") continue;
            if (n < 3) out[n++] = atof(t[i].c_str());
        }
        return n;
    }
};

inline std::vector<std::string> tokenize(const std::string& s)
{
    std::vector<std::string> out;
    size_t i = 0, n = s.size();
    while (i < n)
    {
        char c = s[i];
        if (isspace((unsigned char)c)) { i++; continue; }
        if (c == '/' && i + 1 < n && s[i + 1] == '/')
        {
            while (i < n && s[i] != '\n') i++;
            continue;
        }
        if (c == '/' && i + 1 < n && s[i + 1] == '*')
        {
            i += 2;
            while (i + 1 < n && !(s[i] == '*' && s[i + 1] == '/')) i++;
            i += 2;
            continue;
        }
        if (c == '{' || c == '}' || c == ';')
        {
            out.emplace_back(1, c);
            i++;
            continue;
        }
        if (c == '(' || c == ')')
        {
            // parentheses inside a word (e.g. div(phi,U), grad(U)) stay part of the word: a bare
            // '(' or ')' surrounded by whitespace/number context is a list delimiter
            out.emplace_back(1, c);
            i++;
            continue;
        }
        if (c == '"')
        {
            size_t j = i + 1;
            while (j < n && s[j] != '"') j++;
            out.push_back(s.substr(i + 1, j - i - 1));
            i = j + 1;
            continue;
        }
        // word: may contain balanced parentheses, e.g. div((nuEff*dev2(T(grad(U)))))
        size_t j = i;
        int depth = 0;
        while (j < n)
        {
            char d = s[j];
            if (d == '(')
            {
                // `6(a b c)`: OpenFOAM writes short lists with the size glued to the bracket
                bool count = depth == 0 && j > i;
                for (size_t q = i; count && q < j; q++) count = isdigit((unsigned char)s[q]) != 0;
                if (count) break;
                depth++;
            }
            else if (d == ')')
            {
                if (depth == 0) break;
                depth--;
            }
            else if (depth == 0 && (isspace((unsigned char)d) || d == ';' || d == '{' || d == '}')) break;
            j++;
        }
        out.push_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}

inline void parseDictBody(const std::vector<std::string>& t, size_t& i, Dict& d, bool top)
{
    while (i < t.size())
    {
        if (t[i] == "}")
        {
            if (top) throw Error("unbalanced } in dictionary");
            i++;
            return;
        }
        std::string key = t[i++];
        if (i < t.size() && t[i] == "{")
        {
            i++;
            Dict sub;
            parseDictBody(t, i, sub, false);
            d.subs[key] = sub;
            d.order.push_back(key);
            continue;
        }
        std::vector<std::string> val;
        while (i < t.size() && t[i] != ";")
        {
            if (t[i] == "{")
            {
                // inline sub-dictionary after tokens (e.g. "ddtSchemes { default steadyState; }")
                i++;
                Dict sub;
                parseDictBody(t, i, sub, false);
                d.subs[key] = sub;
                d.order.push_back(key);
                val.clear();
                goto next;
            }
            val.push_back(t[i++]);
        }
        i++; // ';'
        d.entries[key] = val;
    next:;
    }
    if (!top) throw Error("unterminated sub-dictionary");
}

inline Dict parseDict(const std::string& text)
{
    auto t = tokenize(text);
    Dict d;
    size_t i = 0;
    parseDictBody(t, i, d, true);
    return d;
}

inline Dict readDict(const std::string& path) { return parseDict(readFile(path)); }

// ------------------------------------------------------------------------------------------------
// polyMesh lists
// ------------------------------------------------------------------------------------------------
struct ListCursor
{
    const std::string& s;
    size_t i;
    bool binary;
    explicit ListCursor(const std::string& str) : s(str), i(0), binary(false)
    {
        // skip FoamFile header, remember format
        size_t h = s.find("FoamFile");
        if (h != std::string::npos)
        {
            size_t e = s.find('}', h);
            if (e == std::string::npos) throw Error("bad FoamFile header");
            std::string hdr = s.substr(h, e - h);
            binary = hdr.find("binary") != std::string::npos;
            i = e + 1;
        }
    }
    void skipWsComments()
    {
        for (;;)
        {
            while (i < s.size() && isspace((unsigned char)s[i])) i++;
            if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/')
            {
                while (i < s.size() && s[i] != '\n') i++;
                continue;
            }
            if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*')
            {
                i += 2;
                while (i + 1 < s.size() && !(s[i] == '*' && s[i + 1] == '/')) i++;
                i += 2;
                continue;
            }
            break;
        }
    }
    long readInt()
    {
        skipWsComments();
        char* e;
        long v = strtol(s.c_str() + i, &e, 10);
        if (e == s.c_str() + i) throw Error("expected integer in list");
        i = (size_t)(e - s.c_str());
        return v;
    }
    double readDouble()
    {
        skipWsComments();
        char* e;
        double v = strtod(s.c_str() + i, &e);
        if (e == s.c_str() + i) throw Error("expected number in list");
        i = (size_t)(e - s.c_str());
        return v;
    }
    void expect(char c)
    {
        skipWsComments();
        if (i >= s.size() || s[i] != c) throw Error(std::string("expected '") + c + "' in list");
        i++;
    }
    void raw(void* dst, size_t bytes)
    {
        if (i + bytes > s.size()) throw Error("binary list truncated");
        memcpy(dst, s.data() + i, bytes);
        i += bytes;
    }
};

inline void readLabelList(const std::string& path, std::vector<int32_t>& out)
{
    std::string s = readFile(path);
    ListCursor c(s);
    long n = c.readInt();
    out.resize((size_t)n);
    c.expect('(');
    if (c.binary) c.raw(out.data(), (size_t)n * 4);
    else
        for (long k = 0; k < n; k++) out[(size_t)k] = (int32_t)c.readInt();
    c.expect(')');
}

inline void readVectorField(const std::string& path, std::vector<double>& out)
{
    std::string s = readFile(path);
    ListCursor c(s);
    long n = c.readInt();
    out.resize((size_t)n * 3);
    c.expect('(');
    if (c.binary) c.raw(out.data(), (size_t)n * 24);
    else
        for (long k = 0; k < n; k++)
        {
            c.expect('(');
            out[3 * k] = c.readDouble();
            out[3 * k + 1] = c.readDouble();
            out[3 * k + 2] = c.readDouble();
            c.expect(')');
        }
    c.expect(')');
}

inline void readFaceList(const std::string& path, std::vector<int32_t>& off, std::vector<int32_t>& lab)
{
    std::string s = readFile(path);
    ListCursor c(s);
    if (c.binary)
    {
        // faceCompactList: offsets then labels
        long n1 = c.readInt();
        off.resize((size_t)n1);
        c.expect('(');
        c.raw(off.data(), (size_t)n1 * 4);
        c.expect(')');
        long n2 = c.readInt();
        lab.resize((size_t)n2);
        c.expect('(');
        c.raw(lab.data(), (size_t)n2 * 4);
        c.expect(')');
        return;
    }
    long n = c.readInt();
    off.assign(1, 0);
    off.reserve((size_t)n + 1);
    lab.reserve((size_t)n * 4);
    c.expect('(');
    for (long k = 0; k < n; k++)
    {
        long m = c.readInt();
        c.expect('(');
        for (long j = 0; j < m; j++) lab.push_back((int32_t)c.readInt());
        c.expect(')');
        off.push_back((int32_t)lab.size());
    }
    c.expect(')');
}

struct PatchDef
{
    std::string name, type;
    int start, size;
    // cyclic patches (OpenFOAM cyclicPolyPatch entries of constant/polyMesh/boundary)
    std::string neighbourPatch, transform; // transform: rotational | translational | (empty / unknown / noOrdering: from the geometry)
    double rotationAxis[3] = {0, 0, 0}, rotationCentre[3] = {0, 0, 0}, separationVector[3] = {0, 0, 0};
    bool hasAxis = false, hasSeparation = false;
};

inline std::vector<PatchDef> readBoundary(const std::string& path)
{
    std::string s = readFile(path);
    // strip header, then "N ( name { ... } ... )"
    size_t h = s.find("FoamFile");
    size_t i = 0;
    if (h != std::string::npos) i = s.find('}', h) + 1;
    std::string body = s.substr(i);
    size_t lp = body.find('(');
    size_t rp = body.rfind(')');
    if (lp == std::string::npos || rp == std::string::npos) throw Error("bad boundary file");
    Dict d = parseDict(body.substr(lp + 1, rp - lp - 1));
    std::vector<PatchDef> out;
    for (const auto& name : d.order)
    {
        const Dict& p = d.sub(name);
        PatchDef pd;
        pd.name = name;
        pd.type = p.word("type");
        pd.start = atoi(p.word("startFace").c_str());
        pd.size = atoi(p.word("nFaces").c_str());
        if (pd.type == "cyclic")
        {
            pd.neighbourPatch = p.word("neighbourPatch");
            pd.transform = p.wordOr("transform", "");
            if (p.has("rotationAxis")) { p.uniform("rotationAxis", pd.rotationAxis); pd.hasAxis = true; }
            if (p.has("rotationCentre")) p.uniform("rotationCentre", pd.rotationCentre);
            if (p.has("separationVector")) { p.uniform("separationVector", pd.separationVector); pd.hasSeparation = true; }
        }
        out.push_back(pd);
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// JSON (DAOPTION dict)
// ------------------------------------------------------------------------------------------------
struct JVal
{
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;

    const JVal* get(const std::string& k) const
    {
        for (const auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
    double numOr(const std::string& k, double d) const
    {
        const JVal* v = get(k);
        if (!v) return d;
        if (v->kind == Num) return v->num;
        if (v->kind == Bool) return v->b ? 1.0 : 0.0;
        return d;
    }
    std::string strOr(const std::string& k, const std::string& d) const
    {
        const JVal* v = get(k);
        return (v && v->kind == Str) ? v->str : d;
    }
};

struct JParser
{
    const std::string& s;
    size_t i = 0;
    explicit JParser(const std::string& str) : s(str) {}
    void ws()
    {
        while (i < s.size() && isspace((unsigned char)s[i])) i++;
    }
    JVal parse()
    {
        ws();
        if (i >= s.size()) throw Error("json: unexpected end");
        char c = s[i];
        JVal v;
        if (c == '{')
        {
            v.kind = JVal::Obj;
            i++;
            ws();
            if (s[i] == '}') { i++; return v; }
            for (;;)
            {
                ws();
                JVal k = parse();
                if (k.kind != JVal::Str) throw Error("json: key must be string");
                ws();
                if (s[i] != ':') throw Error("json: expected :");
                i++;
                JVal val = parse();
                v.obj.emplace_back(k.str, val);
                ws();
                if (s[i] == ',') { i++; continue; }
                if (s[i] == '}') { i++; break; }
                throw Error("json: expected , or }");
            }
            return v;
        }
        if (c == '[')
        {
            v.kind = JVal::Arr;
            i++;
            ws();
            if (s[i] == ']') { i++; return v; }
            for (;;)
            {
                v.arr.push_back(parse());
                ws();
                if (s[i] == ',') { i++; continue; }
                if (s[i] == ']') { i++; break; }
                throw Error("json: expected , or ]");
            }
            return v;
        }
        if (c == '"')
        {
            v.kind = JVal::Str;
            i++;
            while (i < s.size() && s[i] != '"')
            {
                if (s[i] == '\\' && i + 1 < s.size()) i++;
                v.str.push_back(s[i++]);
            }
            i++;
            return v;
        }
        if (!strncmp(s.c_str() + i, "true", 4)) { v.kind = JVal::Bool; v.b = true; i += 4; return v; }
        if (!strncmp(s.c_str() + i, "false", 5)) { v.kind = JVal::Bool; v.b = false; i += 5; return v; }
        if (!strncmp(s.c_str() + i, "null", 4)) { i += 4; return v; }
        char* e;
        v.num = strtod(s.c_str() + i, &e);
        if (e == s.c_str() + i) throw Error("json: bad token");
        v.kind = JVal::Num;
        i = (size_t)(e - s.c_str());
        return v;
    }
};

inline JVal parseJson(const std::string& s)
{
    JParser p(s);
    return p.parse();
}

} // namespace dab
