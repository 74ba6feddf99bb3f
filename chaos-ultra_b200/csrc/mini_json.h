/*
 * mini_json.h -- just enough JSON for the Newton modules' custom parameters
 * ({"coefficients":[..], "roots":[[re,im],..], "colorMagnifier": n}); the reference uses Gson for this
 * (modules/ModuleNewtonGeneric.java:53-65, util/JsonHelpers.java).  Objects, arrays, numbers, strings, literals.
 */
#ifndef CHAOS_MINI_JSON_H
#define CHAOS_MINI_JSON_H

#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

struct mj_value {
    enum kind_t { NUL, NUM, STR, ARR, OBJ, BOOL } kind = NUL;
    double num = 0;
    std::string str;
    std::vector<mj_value> arr;
    std::map<std::string, mj_value> obj;
    const mj_value *get(const char *key) const
    {
        auto it = obj.find(key);
        return it == obj.end() ? nullptr : &it->second;
    }
};

#define MJ_MAX_DEPTH 64   /* user text like "[[[[..." must not overflow the host stack */
struct mj_parser {
    const char *p;
    bool ok = true;
    int depth = 0;
    explicit mj_parser(const char *text) : p(text ? text : "") {}
    void ws() { while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p; }
    bool eat(char c) { ws(); if (*p == c) { ++p; return true; } return false; }
    mj_value parse_value()
    {
        mj_value v;
        ws();
        if ((*p == '{' || *p == '[') && depth >= MJ_MAX_DEPTH) { ok = false; return v; }
        struct nest { int &d; explicit nest(int &x) : d(x) { ++d; } ~nest() { --d; } } guard(depth);
        if (*p == '{') {
            ++p; v.kind = mj_value::OBJ;
            if (eat('}')) return v;
            do {
                ws();
                mj_value k = parse_value();
                if (k.kind != mj_value::STR || !eat(':')) { ok = false; return v; }
                v.obj[k.str] = parse_value();
            } while (ok && eat(','));
            if (!eat('}')) ok = false;
        } else if (*p == '[') {
            ++p; v.kind = mj_value::ARR;
            if (eat(']')) return v;
            do { v.arr.push_back(parse_value()); } while (ok && eat(','));
            if (!eat(']')) ok = false;
        } else if (*p == '"') {
            ++p; v.kind = mj_value::STR;
            while (*p && *p != '"') { if (*p == '\\' && p[1]) ++p; v.str.push_back(*p++); }
            if (*p == '"') ++p; else ok = false;
        } else if (!strncmp(p, "true", 4)) { p += 4; v.kind = mj_value::BOOL; v.num = 1; }
        else if (!strncmp(p, "false", 5)) { p += 5; v.kind = mj_value::BOOL; }
        else if (!strncmp(p, "null", 4)) { p += 4; }
        else {
            char *end = nullptr;
            v.num = strtod(p, &end);
            if (end == p) ok = false; else { v.kind = mj_value::NUM; p = end; }
        }
        return v;
    }
    bool parse(mj_value &out)
    {
        out = parse_value();
        ws();
        return ok && *p == '\0';
    }
};

#endif
