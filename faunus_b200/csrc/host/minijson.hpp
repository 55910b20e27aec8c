// Minimal JSON value / parser / writer used by the C++ host layer and by the CPU oracle.
//
// The reference consumes one JSON document (src/faunus.cpp:88-135) through nlohmann::json,
// which is not available in this image. This is a small self-contained replacement that
// covers what the hot-path configuration needs: objects, arrays, strings, numbers, bools, null.
// It is a neutral utility (not part of the energy path).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fb {

class Json
{
  public:
    enum class Type
    {
        Null,
        Bool,
        Number,
        String,
        Array,
        Object
    };
    using Array = std::vector<Json>;
    using Member = std::pair<std::string, Json>;
    using Object = std::vector<Member>; // insertion ordered

  private:
    Type type_ = Type::Null;
    bool bool_ = false;
    double num_ = 0.0;
    bool is_int_ = false;
    std::string str_;
    Array arr_;
    Object obj_;

  public:
    Json() = default;
    Json(std::nullptr_t) {}
    Json(bool b)
        : type_(Type::Bool)
        , bool_(b)
    {
    }
    Json(double d)
        : type_(Type::Number)
        , num_(d)
    {
    }
    Json(int i)
        : type_(Type::Number)
        , num_(i)
        , is_int_(true)
    {
    }
    Json(long i)
        : type_(Type::Number)
        , num_(static_cast<double>(i))
        , is_int_(true)
    {
    }
    Json(size_t i)
        : type_(Type::Number)
        , num_(static_cast<double>(i))
        , is_int_(true)
    {
    }
    Json(const char* s)
        : type_(Type::String)
        , str_(s)
    {
    }
    Json(const std::string& s)
        : type_(Type::String)
        , str_(s)
    {
    }
    static Json array()
    {
        Json j;
        j.type_ = Type::Array;
        return j;
    }
    static Json object()
    {
        Json j;
        j.type_ = Type::Object;
        return j;
    }
    template <typename T> static Json fromVector(const std::vector<T>& v)
    {
        Json j = array();
        for (const auto& x : v) {
            j.arr_.emplace_back(x);
        }
        return j;
    }

    Type type() const { return type_; }
    bool is_null() const { return type_ == Type::Null; }
    bool is_bool() const { return type_ == Type::Bool; }
    bool is_number() const { return type_ == Type::Number; }
    bool is_string() const { return type_ == Type::String; }
    bool is_array() const { return type_ == Type::Array; }
    bool is_object() const { return type_ == Type::Object; }

    size_t size() const
    {
        if (is_array()) {
            return arr_.size();
        }
        if (is_object()) {
            return obj_.size();
        }
        return is_null() ? 0 : 1;
    }
    bool empty() const { return size() == 0; }

    // --- typed access (throws on mismatch; numbers accept "inf"/"-inf" strings as nlohmann users do) ---
    double number() const
    {
        if (is_number()) {
            return num_;
        }
        if (is_string()) {
            if (str_ == "inf" || str_ == "oo" || str_ == "infinity") {
                return INFINITY;
            }
            if (str_ == "-inf" || str_ == "-oo" || str_ == "-infinity") {
                return -INFINITY;
            }
        }
        throw std::runtime_error("json: number expected, got " + dump());
    }
    int integer() const { return static_cast<int>(std::llround(number())); }
    bool boolean() const
    {
        if (is_bool()) {
            return bool_;
        }
        if (is_number()) {
            return num_ != 0.0;
        }
        throw std::runtime_error("json: bool expected, got " + dump());
    }
    const std::string& string() const
    {
        if (!is_string()) {
            throw std::runtime_error("json: string expected, got " + dump());
        }
        return str_;
    }
    const Array& items() const
    {
        if (!is_array()) {
            throw std::runtime_error("json: array expected, got " + dump());
        }
        return arr_;
    }
    Array& items()
    {
        if (is_null()) {
            type_ = Type::Array;
        }
        if (!is_array()) {
            throw std::runtime_error("json: array expected");
        }
        return arr_;
    }
    const Object& members() const
    {
        if (!is_object()) {
            throw std::runtime_error("json: object expected, got " + dump());
        }
        return obj_;
    }

    // --- object access ---
    bool contains(const std::string& key) const { return find(key) != nullptr; }
    const Json* find(const std::string& key) const
    {
        if (!is_object()) {
            return nullptr;
        }
        for (const auto& m : obj_) {
            if (m.first == key) {
                return &m.second;
            }
        }
        return nullptr;
    }
    const Json& at(const std::string& key) const
    {
        if (const auto* p = find(key)) {
            return *p;
        }
        throw std::runtime_error("json: missing key '" + key + "'");
    }
    Json& operator[](const std::string& key)
    {
        if (is_null()) {
            type_ = Type::Object;
        }
        if (!is_object()) {
            throw std::runtime_error("json: object expected for key '" + key + "'");
        }
        for (auto& m : obj_) {
            if (m.first == key) {
                return m.second;
            }
        }
        obj_.emplace_back(key, Json());
        return obj_.back().second;
    }
    const Json& at(size_t i) const { return items().at(i); }
    void push_back(const Json& j) { items().push_back(j); }

    double value(const std::string& key, double fallback) const
    {
        const auto* p = find(key);
        return p ? p->number() : fallback;
    }
    int value(const std::string& key, int fallback) const
    {
        const auto* p = find(key);
        return p ? p->integer() : fallback;
    }
    bool value(const std::string& key, bool fallback) const
    {
        const auto* p = find(key);
        return p ? p->boolean() : fallback;
    }
    std::string value(const std::string& key, const std::string& fallback) const
    {
        const auto* p = find(key);
        return p ? p->string() : fallback;
    }
    std::string value(const std::string& key, const char* fallback) const
    {
        return value(key, std::string(fallback));
    }
    std::vector<double> numbers() const
    {
        std::vector<double> v;
        for (const auto& x : items()) {
            v.push_back(x.number());
        }
        return v;
    }

    /** Single `{key: value}` object as used all over the Faunus input (src/aux/json_support.h) */
    const Member& single() const
    {
        if (!is_object() || obj_.size() != 1) {
            throw std::runtime_error("json: single-key object expected, got " + dump());
        }
        return obj_.front();
    }

    // --- serialisation ---
    std::string dump() const
    {
        std::string out;
        write(out);
        return out;
    }

    static Json parse(const std::string& text)
    {
        Parser p{text.c_str(), text.c_str() + text.size()};
        Json j = p.value();
        p.skip();
        if (p.cur != p.end) {
            throw std::runtime_error("json: trailing characters");
        }
        return j;
    }

    // --- Universal Binary JSON (the reference's `.ubj` state files: nlohmann `json::to_ubjson` with its defaults —
    // no container size / type optimisation — src/analysis.cpp:656-668; read back by `json::from_ubjson`,
    // src/faunus.cpp:435-449). Integers take the smallest of i, U, I, l, L that holds them, every other number is a
    // big-endian 'D'; strings and keys carry their length as such an integer; the reader also accepts the optimised
    // containers ('$' type, '#' count) and 'd', 'C' that other writers produce. ---
    std::string toUbjson() const
    {
        std::string out;
        writeUbjson(out);
        return out;
    }

    static Json fromUbjson(const std::string& bytes)
    {
        UbjsonReader r{reinterpret_cast<const unsigned char*>(bytes.data()),
                       reinterpret_cast<const unsigned char*>(bytes.data()) + bytes.size()};
        Json j = r.value(r.marker());
        if (r.cur != r.end) {
            throw std::runtime_error("ubjson: trailing bytes");
        }
        return j;
    }

  private:
    static void ubjsonBigEndian(std::string& out, unsigned long long bits, int n_bytes)
    {
        for (int b = n_bytes - 1; b >= 0; --b) {
            out.push_back(static_cast<char>((bits >> (8 * b)) & 0xffu));
        }
    }
    static void ubjsonInteger(std::string& out, long long v)
    {
        if (v >= -128 && v <= 127) {
            out.push_back('i');
            ubjsonBigEndian(out, static_cast<unsigned long long>(v), 1);
        }
        else if (v >= 0 && v <= 255) {
            out.push_back('U');
            ubjsonBigEndian(out, static_cast<unsigned long long>(v), 1);
        }
        else if (v >= -32768 && v <= 32767) {
            out.push_back('I');
            ubjsonBigEndian(out, static_cast<unsigned long long>(v), 2);
        }
        else if (v >= -2147483648LL && v <= 2147483647LL) {
            out.push_back('l');
            ubjsonBigEndian(out, static_cast<unsigned long long>(v), 4);
        }
        else {
            out.push_back('L');
            ubjsonBigEndian(out, static_cast<unsigned long long>(v), 8);
        }
    }
    void writeUbjson(std::string& out) const
    {
        switch (type_) {
        case Type::Null:
            out.push_back('Z');
            break;
        case Type::Bool:
            out.push_back(bool_ ? 'T' : 'F');
            break;
        case Type::Number:
            if (is_int_ && std::fabs(num_) < 9.0e18) {
                ubjsonInteger(out, static_cast<long long>(num_));
            }
            else {
                unsigned long long bits;
                static_assert(sizeof(bits) == sizeof(num_));
                std::memcpy(&bits, &num_, sizeof(bits));
                out.push_back('D');
                ubjsonBigEndian(out, bits, 8);
            }
            break;
        case Type::String:
            out.push_back('S');
            ubjsonInteger(out, static_cast<long long>(str_.size()));
            out += str_;
            break;
        case Type::Array:
            out.push_back('[');
            for (const auto& x : arr_) {
                x.writeUbjson(out);
            }
            out.push_back(']');
            break;
        case Type::Object:
            out.push_back('{');
            for (const auto& [key, x] : obj_) {
                ubjsonInteger(out, static_cast<long long>(key.size()));
                out += key;
                x.writeUbjson(out);
            }
            out.push_back('}');
            break;
        }
    }

    struct UbjsonReader
    {
        const unsigned char* cur;
        const unsigned char* end;
        [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("ubjson: ") + what); }
        unsigned char marker()
        {
            if (cur >= end) {
                fail("unexpected end of input");
            }
            return *cur++;
        }
        unsigned long long bigEndian(int n_bytes)
        {
            if (end - cur < n_bytes) {
                fail("unexpected end of input");
            }
            unsigned long long bits = 0;
            for (int b = 0; b < n_bytes; ++b) {
                bits = (bits << 8) | *cur++;
            }
            return bits;
        }
        bool integerOf(unsigned char m, long long& v)
        {
            switch (m) {
            case 'i':
                v = static_cast<signed char>(bigEndian(1));
                return true;
            case 'U':
                v = static_cast<long long>(bigEndian(1));
                return true;
            case 'I':
                v = static_cast<short>(bigEndian(2));
                return true;
            case 'l':
                v = static_cast<int>(bigEndian(4));
                return true;
            case 'L':
                v = static_cast<long long>(bigEndian(8));
                return true;
            default:
                return false;
            }
        }
        size_t length()
        {
            long long n = 0;
            if (!integerOf(marker(), n) || n < 0) {
                fail("length expected");
            }
            if (n > static_cast<long long>(end - cur)) {
                fail("length beyond the input");
            }
            return static_cast<size_t>(n);
        }
        std::string text()
        {
            const size_t n = length();
            if (static_cast<size_t>(end - cur) < n) {
                fail("unexpected end of input");
            }
            std::string s(reinterpret_cast<const char*>(cur), n);
            cur += n;
            return s;
        }
        /** optional "$ type" and "# count" after '[' or '{' */
        void containerHeader(unsigned char& type, long long& count)
        {
            type = 0;
            count = -1;
            if (cur < end && *cur == '$') {
                ++cur;
                type = marker();
                if (cur >= end || *cur != '#') {
                    fail("a typed container needs a count");
                }
            }
            if (cur < end && *cur == '#') {
                ++cur;
                count = static_cast<long long>(length());
            }
        }
        Json value(unsigned char m)
        {
            long long v = 0;
            if (integerOf(m, v)) {
                Json j(static_cast<double>(v));
                j.is_int_ = true;
                return j;
            }
            switch (m) {
            case 'Z':
                return Json();
            case 'T':
                return Json(true);
            case 'F':
                return Json(false);
            case 'D': {
                const unsigned long long bits = bigEndian(8);
                double d;
                std::memcpy(&d, &bits, sizeof(d));
                return Json(d);
            }
            case 'd': {
                const unsigned bits = static_cast<unsigned>(bigEndian(4));
                float f;
                std::memcpy(&f, &bits, sizeof(f));
                return Json(static_cast<double>(f));
            }
            case 'C':
                return Json(std::string(1, static_cast<char>(bigEndian(1))));
            case 'S':
                return Json(text());
            case '[': {
                Json j = Json::array();
                unsigned char type;
                long long count;
                containerHeader(type, count);
                if (count >= 0) {
                    for (long long i = 0; i < count; ++i) {
                        j.arr_.push_back(value(type ? type : marker()));
                    }
                }
                else {
                    for (unsigned char e = marker(); e != ']'; e = marker()) {
                        j.arr_.push_back(value(e));
                    }
                }
                return j;
            }
            case '{': {
                Json j = Json::object();
                unsigned char type;
                long long count;
                containerHeader(type, count);
                for (long long i = 0; count < 0 || i < count; ++i) {
                    if (count < 0 && cur < end && *cur == '}') {
                        ++cur;
                        break;
                    }
                    std::string key = text();
                    j.obj_.emplace_back(std::move(key), value(type ? type : marker()));
                }
                return j;
            }
            default:
                fail("unknown type marker");
            }
        }
    };

    static void writeString(std::string& out, const std::string& s)
    {
        out.push_back('"');
        for (unsigned char c : s) {
            switch (c) {
            case '"':
                out += "\\\"";
                break;
            case '\\':
                out += "\\\\";
                break;
            case '\n':
                out += "\\n";
                break;
            case '\t':
                out += "\\t";
                break;
            case '\r':
                out += "\\r";
                break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    std::snprintf(buf, sizeof buf, "\\u%04x", c);
                    out += buf;
                }
                else {
                    out.push_back(static_cast<char>(c));
                }
            }
        }
        out.push_back('"');
    }

    void write(std::string& out) const
    {
        switch (type_) {
        case Type::Null:
            out += "null";
            break;
        case Type::Bool:
            out += bool_ ? "true" : "false";
            break;
        case Type::Number: {
            char buf[40];
            if (std::isnan(num_)) {
                out += "null";
            }
            else if (std::isinf(num_)) {
                out += num_ > 0 ? "\"inf\"" : "\"-inf\"";
            }
            else if (is_int_ || (std::floor(num_) == num_ && std::fabs(num_) < 1e15)) {
                std::snprintf(buf, sizeof buf, "%.0f", num_);
                out += buf;
                if (!is_int_) {
                    out += ".0";
                }
            }
            else {
                std::snprintf(buf, sizeof buf, "%.17g", num_);
                out += buf;
            }
            break;
        }
        case Type::String:
            writeString(out, str_);
            break;
        case Type::Array: {
            out.push_back('[');
            bool first = true;
            for (const auto& x : arr_) {
                if (!first) {
                    out.push_back(',');
                }
                first = false;
                x.write(out);
            }
            out.push_back(']');
            break;
        }
        case Type::Object: {
            out.push_back('{');
            bool first = true;
            for (const auto& m : obj_) {
                if (!first) {
                    out.push_back(',');
                }
                first = false;
                writeString(out, m.first);
                out.push_back(':');
                m.second.write(out);
            }
            out.push_back('}');
            break;
        }
        }
    }

    struct Parser
    {
        const char* cur;
        const char* end;

        void skip()
        {
            while (cur < end && (*cur == ' ' || *cur == '\n' || *cur == '\t' || *cur == '\r')) {
                ++cur;
            }
        }
        [[noreturn]] void fail(const char* what) const
        {
            throw std::runtime_error(std::string("json parse error: ") + what);
        }
        bool match(const char* lit)
        {
            const size_t n = std::strlen(lit);
            if (static_cast<size_t>(end - cur) >= n && std::strncmp(cur, lit, n) == 0) {
                cur += n;
                return true;
            }
            return false;
        }
        static void appendUtf8(std::string& s, unsigned cp)
        {
            if (cp < 0x80) {
                s.push_back(static_cast<char>(cp));
            }
            else if (cp < 0x800) {
                s.push_back(static_cast<char>(0xC0 | (cp >> 6)));
                s.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
            }
            else {
                s.push_back(static_cast<char>(0xE0 | (cp >> 12)));
                s.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
                s.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
            }
        }
        std::string parseString()
        {
            if (*cur != '"') {
                fail("string expected");
            }
            ++cur;
            std::string s;
            while (cur < end && *cur != '"') {
                if (*cur == '\\') {
                    ++cur;
                    if (cur >= end) {
                        fail("bad escape");
                    }
                    switch (*cur) {
                    case 'n':
                        s.push_back('\n');
                        break;
                    case 't':
                        s.push_back('\t');
                        break;
                    case 'r':
                        s.push_back('\r');
                        break;
                    case 'b':
                        s.push_back('\b');
                        break;
                    case 'f':
                        s.push_back('\f');
                        break;
                    case 'u': {
                        if (end - cur < 5) {
                            fail("bad unicode escape");
                        }
                        char hex[5] = {cur[1], cur[2], cur[3], cur[4], 0};
                        appendUtf8(s, static_cast<unsigned>(std::strtoul(hex, nullptr, 16)));
                        cur += 4;
                        break;
                    }
                    default:
                        s.push_back(*cur);
                    }
                    ++cur;
                }
                else {
                    s.push_back(*cur++);
                }
            }
            if (cur >= end) {
                fail("unterminated string");
            }
            ++cur;
            return s;
        }
        Json value()
        {
            skip();
            if (cur >= end) {
                fail("unexpected end");
            }
            if (*cur == '{') {
                ++cur;
                Json j = Json::object();
                skip();
                if (cur < end && *cur == '}') {
                    ++cur;
                    return j;
                }
                while (true) {
                    skip();
                    std::string key = parseString();
                    skip();
                    if (cur >= end || *cur != ':') {
                        fail("':' expected");
                    }
                    ++cur;
                    j.obj_.emplace_back(std::move(key), value());
                    skip();
                    if (cur < end && *cur == ',') {
                        ++cur;
                        continue;
                    }
                    if (cur < end && *cur == '}') {
                        ++cur;
                        return j;
                    }
                    fail("',' or '}' expected");
                }
            }
            if (*cur == '[') {
                ++cur;
                Json j = Json::array();
                skip();
                if (cur < end && *cur == ']') {
                    ++cur;
                    return j;
                }
                while (true) {
                    j.arr_.push_back(value());
                    skip();
                    if (cur < end && *cur == ',') {
                        ++cur;
                        continue;
                    }
                    if (cur < end && *cur == ']') {
                        ++cur;
                        return j;
                    }
                    fail("',' or ']' expected");
                }
            }
            if (*cur == '"') {
                return Json(parseString());
            }
            if (match("true")) {
                return Json(true);
            }
            if (match("false")) {
                return Json(false);
            }
            if (match("null")) {
                return Json();
            }
            if (match("NaN")) {
                return Json(std::nan(""));
            }
            if (match("Infinity")) {
                return Json(static_cast<double>(INFINITY));
            }
            if (match("-Infinity")) {
                return Json(-static_cast<double>(INFINITY));
            }
            // number
            char* stop = nullptr;
            const double d = std::strtod(cur, &stop);
            if (stop == cur) {
                fail("value expected");
            }
            bool is_int = true;
            for (const char* p = cur; p < stop; ++p) {
                if (*p == '.' || *p == 'e' || *p == 'E') {
                    is_int = false;
                }
            }
            cur = stop;
            Json j(d);
            j.is_int_ = is_int;
            return j;
        }
    };
};

} // namespace fb
