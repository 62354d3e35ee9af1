// frontend.cpp -- Lua-subset interpreter + Opt DSL + CUDA code generation.  See frontend.h.
//
// DSL surface mirrored (names and argument order): API/src/o.t Dim :320-324, Unknown/Image/Array :930-1006,
// ComputedArray :1007-1040, Graph :1043-1053, Param :1055-1058, image call syntax :1061-1103, InBounds /
// InBoundsExpanded :1100-1115, Exclude :2462-2465, UsePreconditioner :942-944; API/src/lib.t Energy :6,
// All :13, And/Or/Not :31-33, Matrix3x3Mul :40, Dot3 :47, normalize :55, length :59, Rotate3D :77, Rotate2D :92,
// L_2_norm :99, Select :116, Stencil :117; vector lifting of operators API/src/ad.t:330-348; default bounds rule
// for centred residuals o.t:1930-1934; residual domain classification o.t:1895-1936.
#include "frontend.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace optb {
namespace fe {
namespace {

[[noreturn]] void fail(const std::string& msg, int line = 0) {
    throw std::runtime_error(line > 0 ? ("line " + std::to_string(line) + ": " + msg) : msg);
}

// ===================================================================================== expression DAG
enum class Op { Const, Param, Load, Unknown, Index, Bounds, Add, Sub, Mul, Div, Neg, Pow, Sin, Cos, Sqrt, Abs, Exp, Log,
                Tan, Tanh, Sinh, Cosh, Asin, Acos, Atan, Asinh, Acosh, Log10,   // the rest of ad.t:747-763 (single-argument ones)
                Less, LessEq, Greater, GreaterEq, Eq, Or, Not, Select, Sample };
// single-argument math functions beyond the common ones: DSL name, device function (generic_device.cuh), host folding
struct Math1 { Op op; const char* name; const char* device; double (*host)(double); };
static const Math1 kMath1[] = {
    { Op::Tan, "tan", "geTan", [](double x) { return std::tan(x); } },     { Op::Tanh, "tanh", "geTanh", [](double x) { return std::tanh(x); } },
    { Op::Sinh, "sinh", "geSinh", [](double x) { return std::sinh(x); } },  { Op::Cosh, "cosh", "geCosh", [](double x) { return std::cosh(x); } },
    { Op::Asin, "asin", "geAsin", [](double x) { return std::asin(x); } },  { Op::Acos, "acos", "geAcos", [](double x) { return std::acos(x); } },
    { Op::Atan, "atan", "geAtan", [](double x) { return std::atan(x); } },  { Op::Asinh, "asinh", "geAsinh", [](double x) { return std::asinh(x); } },
    { Op::Acosh, "acosh", "geAcosh", [](double x) { return std::acosh(x); } }, { Op::Log10, "log10", "geLog10", [](double x) { return std::log10(x); } },
};
static const Math1* math1Of(Op op) { for (const Math1& m : kMath1) if (m.op == op) return &m; return nullptr; }
struct Node {
    Op op = Op::Const;
    double value = 0;        // Const
    int id = -1;             // Param: params[] index; Load/Unknown: images[] index; Index: dimension
    bool graph = false;      // Load/Unknown: accessed through a graph element
    int off[3] = { 0, 0, 0 };   // Load/Unknown: offsets (centred);  Index: off[0] = shift;  Bounds: min
    int off2[3] = { 0, 0, 0 };  // Bounds: max
    int slot = -1;           // Load/Unknown via graph: global graph-slot id
    int chan = 0;
    int nd = 0;              // Bounds: number of dims
    // Sample (bilinear read of a 2-D image at computed coordinates, o.t:582-589): id = image, off[0]/off[1] = images holding d/dx and d/dy
    // (-1: none), chan; args = { x, y }
    std::vector<int> args;
    int height = 0;          // longest path to a leaf; bounded, because analysis and code generation walk the DAG recursively
};
struct Dag {
    std::vector<Node> nodes;
    std::unordered_map<std::string, int> index;
    static constexpr int kMaxHeight = 4000, kMaxNodes = 2000000;
    int add(const Node& n) {
        std::ostringstream k;
        k.precision(17);
        k << (int)n.op << '|' << n.value << '|' << n.id << '|' << n.graph << '|' << n.off[0] << ',' << n.off[1] << ',' << n.off[2] << '|' << n.off2[0] << ',' << n.off2[1] << ','
          << n.off2[2] << '|' << n.slot << '|' << n.chan << '|' << n.nd;
        for (int a : n.args) k << ':' << a;
        auto it = index.find(k.str());
        if (it != index.end()) return it->second;
        int h = 0;
        for (int a : n.args) h = std::max(h, nodes[a].height + 1);
        if (h > kMaxHeight) fail("residual expression too deep (a chain of more than " + std::to_string(kMaxHeight) + " operations)");
        if ((int)nodes.size() >= kMaxNodes) fail("residual expressions too large (more than " + std::to_string(kMaxNodes) + " distinct operations)");
        nodes.push_back(n);
        nodes.back().height = h;
        index[k.str()] = (int)nodes.size() - 1;
        return (int)nodes.size() - 1;
    }
    int constant(double v) { Node n; n.op = Op::Const; n.value = v; return add(n); }
    bool isConst(int e, double* v = nullptr) const { if (nodes[e].op == Op::Const) { if (v) *v = nodes[e].value; return true; } return false; }
    int unary(Op op, int a) {
        double v;
        if (isConst(a, &v)) {
            switch (op) {
                case Op::Neg: return constant(-v); case Op::Sin: return constant(std::sin(v)); case Op::Cos: return constant(std::cos(v));
                case Op::Sqrt: return constant(std::sqrt(v)); case Op::Abs: return constant(std::fabs(v)); case Op::Exp: return constant(std::exp(v));
                case Op::Log: return constant(std::log(v)); case Op::Not: return constant(v != 0 ? 0.0 : 1.0);
                default: if (const Math1* m = math1Of(op)) return constant(m->host(v)); break;
            }
        }
        Node n; n.op = op; n.args = { a }; return add(n);
    }
    int binary(Op op, int a, int b) {
        double x, y;
        if (isConst(a, &x) && isConst(b, &y)) {
            switch (op) {
                case Op::Add: return constant(x + y); case Op::Sub: return constant(x - y); case Op::Mul: return constant(x * y); case Op::Div: return constant(x / y);
                case Op::Pow: return constant(std::pow(x, y)); case Op::Less: return constant(x < y); case Op::LessEq: return constant(x <= y);
                case Op::Greater: return constant(x > y); case Op::GreaterEq: return constant(x >= y); case Op::Eq: return constant(x == y);
                case Op::Or: return constant((x != 0 || y != 0) ? 1.0 : 0.0); default: break;
            }
        }
        // cheap identities keep the generated code small (x*1, x*0, x+0, x-0); never across a Select
        if (op == Op::Mul) { if (isConst(a, &x) && x == 1) return b; if (isConst(b, &y) && y == 1) return a; if ((isConst(a, &x) && x == 0) || (isConst(b, &y) && y == 0)) return constant(0); }
        if (op == Op::Add) { if (isConst(a, &x) && x == 0) return b; if (isConst(b, &y) && y == 0) return a; }
        if (op == Op::Sub) { if (isConst(b, &y) && y == 0) return a; }
        Node n; n.op = op; n.args = { a, b }; return add(n);
    }
    int select(int c, int a, int b) {
        double v;
        if (isConst(c, &v)) return v != 0 ? a : b;
        if (a == b) return a;
        Node n; n.op = Op::Select; n.args = { c, a, b }; return add(n);
    }
};

// ===================================================================================== DSL objects and Lua values
struct Table; struct Closure; struct Interp;
struct Object { virtual ~Object() {} virtual const char* kind() const = 0; };
struct Value {
    enum Type { Nil, Bool, Num, Str, Tab, Fn, Builtin, Obj } type = Nil;
    bool b = false; double n = 0; std::string s;
    std::shared_ptr<Table> t; std::shared_ptr<Closure> fn;
    std::shared_ptr<std::function<std::vector<Value>(Interp&, std::vector<Value>&, int)>> bf;
    std::shared_ptr<Object> obj;
    static Value num(double v) { Value x; x.type = Num; x.n = v; return x; }
    static Value boolean(bool v) { Value x; x.type = Bool; x.b = v; return x; }
    static Value str(const std::string& v) { Value x; x.type = Str; x.s = v; return x; }
    static Value object(std::shared_ptr<Object> o) { Value x; x.type = Obj; x.obj = std::move(o); return x; }
    bool truthy() const { return !(type == Nil || (type == Bool && !b)); }
};
struct Table {
    std::map<double, Value> arr; std::map<std::string, Value> fields;
    Value get(const Value& k) const {
        if (k.type == Value::Num) { auto it = arr.find(k.n); return it == arr.end() ? Value() : it->second; }
        if (k.type == Value::Str) { auto it = fields.find(k.s); return it == fields.end() ? Value() : it->second; }
        return Value();
    }
    void set(const Value& k, const Value& v) {
        if (k.type == Value::Num) { if (v.type == Value::Nil) arr.erase(k.n); else arr[k.n] = v; }
        else if (k.type == Value::Str) { if (v.type == Value::Nil) fields.erase(k.s); else fields[k.s] = v; }
        else fail("unsupported table key type");
    }
    int length() const { int n = 0; while (arr.count((double)(n + 1))) ++n; return n; }
};

struct ExprObj : Object { int e; explicit ExprObj(int e_) : e(e_) {} const char* kind() const override { return "expression"; } };
struct VecObj : Object { std::vector<int> e; const char* kind() const override { return "vector"; } };
struct DimObj : Object { std::string name; int idx = 0; const char* kind() const override { return "Dim"; } };
struct TypeObj : Object { ScalarType st; int channels; TypeObj(ScalarType s, int c) : st(s), channels(c) {} const char* kind() const override { return "type"; } };
struct ImageObj : Object {
    int id = -1;                                  // images[] index, or -1 for a computed array
    std::vector<int> dims; int channels = 1;
    std::vector<int> computed;                    // per channel defining expression (ComputedArray)
    const char* kind() const override { return "image"; }
};
struct SampledObj : Object { int image = -1, dx = -1, dy = -1; int channels = 1; const char* kind() const override { return "sampled image"; } };
struct GraphObj : Object { int id = -1; std::map<std::string, int> slotOf; const char* kind() const override { return "graph"; } };
struct GraphElemObj : Object { int graph = -1; int slot = -1; int globalSlot = -1; std::vector<int> dims; const char* kind() const override { return "graph element"; } };

// ===================================================================================== Lua-subset AST
struct Expr; struct Stat;
using ExprP = std::shared_ptr<Expr>; using StatP = std::shared_ptr<Stat>;
using Block = std::vector<StatP>;
struct Expr {
    enum K { Nil, True, False, Number, String, Name, Index, Call, Method, Function, Binop, Unop, TableCtor, Vararg } k = Nil;
    int line = 0; double num = 0; std::string str;                       // Number / String / Name / Method name / operator
    int depth = 0;                                                       // height of the tree below this node (bounded by the parser)
    ExprP a, b;                                                          // Index(a,b), Binop(a,b), Unop(a), Call(a), Method(a)
    std::vector<ExprP> args;                                             // Call / Method args; TableCtor positional values
    std::vector<std::pair<ExprP, ExprP>> fields;                         // TableCtor keyed fields
    std::vector<std::string> params; bool vararg = false; std::shared_ptr<Block> body;   // Function
};
struct Stat {
    enum K { Local, Assign, CallStat, Do, While, Repeat, If, NumFor, GenFor, Return, Break } k = Local;
    int line = 0;
    std::vector<std::string> names; std::vector<ExprP> targets, exprs;   // Local / Assign / GenFor names / Return
    ExprP e1, e2, e3;                                                    // NumFor bounds, While/Repeat cond, CallStat
    std::shared_ptr<Block> body; std::vector<std::pair<ExprP, std::shared_ptr<Block>>> clauses; std::shared_ptr<Block> orelse;
};

// ------------------------------------------------------------------------------------- lexer
struct Token { enum T { Eof, Name, Number, String, Sym, Kw } t = Eof; std::string s; double n = 0; int line = 1; };
struct Lexer {
    const std::string& src; size_t i = 0; int line = 1;
    explicit Lexer(const std::string& s) : src(s) { if (s.size() >= 3 && (unsigned char)s[0] == 0xEF && (unsigned char)s[1] == 0xBB && (unsigned char)s[2] == 0xBF) i = 3; }
    static bool kw(const std::string& s) {
        static const char* k[] = { "and", "break", "do", "else", "elseif", "end", "false", "for", "function", "if", "in", "local", "nil", "not", "or", "repeat", "return", "then", "true", "until", "while" };
        for (const char* x : k) if (s == x) return true;
        return false;
    }
    bool longOpen(size_t at, size_t* lvl) const {
        if (at >= src.size() || src[at] != '[') return false;
        size_t k = at + 1, e = 0; while (k < src.size() && src[k] == '=') { ++k; ++e; }
        if (k < src.size() && src[k] == '[') { *lvl = e; return true; }
        return false;
    }
    std::string readLong(size_t lvl) {   // positioned at the opening '['
        i += lvl + 2;
        std::string close = "]" + std::string(lvl, '=') + "]";
        size_t e = src.find(close, i);
        if (e == std::string::npos) fail("unterminated long bracket", line);
        std::string body = src.substr(i, e - i);
        for (char c : body) if (c == '\n') ++line;
        i = e + close.size();
        return body;
    }
    Token next() {
        for (;;) {
            while (i < src.size() && isspace((unsigned char)src[i])) { if (src[i] == '\n') ++line; ++i; }
            if (i + 1 < src.size() && src[i] == '-' && src[i + 1] == '-') {
                size_t lvl;
                if (longOpen(i + 2, &lvl)) { i += 2; readLong(lvl); continue; }
                while (i < src.size() && src[i] != '\n') ++i;
                continue;
            }
            break;
        }
        Token t; t.line = line;
        if (i >= src.size()) return t;
        char c = src[i];
        if (isalpha((unsigned char)c) || c == '_') {
            size_t j = i; while (j < src.size() && (isalnum((unsigned char)src[j]) || src[j] == '_')) ++j;
            t.s = src.substr(i, j - i); t.t = kw(t.s) ? Token::Kw : Token::Name; i = j; return t;
        }
        if (isdigit((unsigned char)c) || (c == '.' && i + 1 < src.size() && isdigit((unsigned char)src[i + 1]))) {
            char* end = nullptr; t.n = strtod(src.c_str() + i, &end); t.t = Token::Number;
            size_t j = end - src.c_str();
            if (j + 1 < src.size() + 1 && j < src.size() && (src[j] == 'f' || src[j] == 'F') ) ++j;   // Terra-style float suffix
            i = j; return t;
        }
        if (c == '"' || c == '\'') {
            size_t j = i + 1; std::string v;
            while (j < src.size() && src[j] != c) {
                if (src[j] == '\\' && j + 1 < src.size()) { ++j; char e = src[j]; v += (e == 'n' ? '\n' : e == 't' ? '\t' : e); }
                else v += src[j];
                ++j;
            }
            if (j >= src.size()) fail("unterminated string", line);
            i = j + 1; t.t = Token::String; t.s = v; return t;
        }
        { size_t lvl; if (longOpen(i, &lvl)) { t.t = Token::String; t.s = readLong(lvl); return t; } }
        static const char* s3[] = { "..." }; static const char* s2[] = { "==", "~=", "<=", ">=", "..", "::" };
        for (const char* o : s3) if (src.compare(i, 3, o) == 0) { t.t = Token::Sym; t.s = o; i += 3; return t; }
        for (const char* o : s2) if (src.compare(i, 2, o) == 0) { t.t = Token::Sym; t.s = o; i += 2; return t; }
        t.t = Token::Sym; t.s = std::string(1, c); ++i; return t;
    }
};

// ------------------------------------------------------------------------------------- parser
struct Parser {
    Lexer lx; Token tok, ahead; bool hasAhead = false;
    // limits that keep the recursive-descent parser, the tree-walking interpreter and the tree's destructor off the end of the stack
    int nest = 0;
    static constexpr int kMaxNest = 200, kMaxDepth = 2000;
    struct NestGuard { Parser& p; explicit NestGuard(Parser& q) : p(q) { if (++p.nest > kMaxNest) fail("the specification nests too deeply (more than " + std::to_string(kMaxNest) + " levels)", p.tok.line); } ~NestGuard() { --p.nest; } };
    ExprP fin(ExprP e) {   // records the height of a freshly built node: long operator / suffix chains are built iteratively, not recursively
        int d = 0;
        if (e->a) d = std::max(d, e->a->depth);
        if (e->b) d = std::max(d, e->b->depth);
        for (auto& x : e->args) if (x) d = std::max(d, x->depth);
        for (auto& f : e->fields) { if (f.first) d = std::max(d, f.first->depth); if (f.second) d = std::max(d, f.second->depth); }
        e->depth = d + 1;
        if (e->depth > kMaxDepth) fail("expression too long (more than " + std::to_string(kMaxDepth) + " chained operators)", e->line);
        return e;
    }
    explicit Parser(const std::string& s) : lx(s) { tok = lx.next(); }
    void advance() { if (hasAhead) { tok = ahead; hasAhead = false; } else tok = lx.next(); }
    const Token& peek() { if (!hasAhead) { ahead = lx.next(); hasAhead = true; } return ahead; }
    bool isSym(const char* s) const { return tok.t == Token::Sym && tok.s == s; }
    bool isKw(const char* s) const { return tok.t == Token::Kw && tok.s == s; }
    bool acceptSym(const char* s) { if (isSym(s)) { advance(); return true; } return false; }
    bool acceptKw(const char* s) { if (isKw(s)) { advance(); return true; } return false; }
    void expectSym(const char* s) { if (!acceptSym(s)) fail(std::string("expected '") + s + "' near '" + tok.s + "'", tok.line); }
    void expectKw(const char* s) { if (!acceptKw(s)) fail(std::string("expected '") + s + "' near '" + tok.s + "'", tok.line); }
    std::string expectName() { if (tok.t != Token::Name) fail("expected a name near '" + tok.s + "'", tok.line); std::string s = tok.s; advance(); return s; }

    bool blockEnd() const { return tok.t == Token::Eof || isKw("end") || isKw("else") || isKw("elseif") || isKw("until"); }
    std::shared_ptr<Block> block() {
        NestGuard guard(*this);
        auto b = std::make_shared<Block>();
        while (!blockEnd()) {
            if (acceptSym(";")) continue;
            StatP s = statement();
            b->push_back(s);
            if (s->k == Stat::Return || s->k == Stat::Break) { acceptSym(";"); break; }
        }
        return b;
    }
    ExprP mk(Expr::K k) { auto e = std::make_shared<Expr>(); e->k = k; e->line = tok.line; return e; }
    ExprP functionBody(bool method) {
        ExprP f = mk(Expr::Function);
        if (method) f->params.push_back("self");
        expectSym("(");
        if (!isSym(")")) {
            do {
                if (acceptSym("...")) { f->vararg = true; break; }
                f->params.push_back(expectName());
            } while (acceptSym(","));
        }
        expectSym(")");
        f->body = block();
        expectKw("end");
        return f;
    }
    StatP statement() {
        auto s = std::make_shared<Stat>(); s->line = tok.line;
        if (acceptKw("local")) {
            if (acceptKw("function")) {
                s->k = Stat::Local; std::string n = expectName(); s->names.push_back(n);
                // `local function f` : f is in scope inside its own body -- emulate by declaring first
                auto decl = std::make_shared<Stat>(); decl->k = Stat::Local; decl->line = s->line; decl->names.push_back(n);
                auto asg = std::make_shared<Stat>(); asg->k = Stat::Assign; asg->line = s->line;
                ExprP nm = mk(Expr::Name); nm->str = n; asg->targets.push_back(nm); asg->exprs.push_back(functionBody(false));
                auto blk = std::make_shared<Stat>(); blk->k = Stat::Do; blk->line = s->line; blk->body = std::make_shared<Block>();
                // cannot wrap in Do (scope): return the two statements via a marker block flagged as "splice"
                blk->names.push_back("__splice__"); blk->body->push_back(decl); blk->body->push_back(asg);
                return blk;
            }
            s->k = Stat::Local;
            do { s->names.push_back(expectName()); } while (acceptSym(","));
            if (acceptSym("=")) s->exprs = exprList();
            return s;
        }
        if (acceptKw("function")) {
            s->k = Stat::Assign;
            ExprP target = mk(Expr::Name); target->str = expectName();
            bool method = false;
            while (isSym(".") || isSym(":")) {
                bool colon = isSym(":"); advance();
                ExprP idx = mk(Expr::Index); idx->a = target; ExprP key = mk(Expr::String); key->str = expectName(); idx->b = key; target = idx;
                if (colon) { method = true; break; }
            }
            s->targets.push_back(target); s->exprs.push_back(functionBody(method));
            return s;
        }
        if (acceptKw("if")) {
            s->k = Stat::If;
            ExprP c = expression(); expectKw("then"); s->clauses.push_back({ c, block() });
            while (acceptKw("elseif")) { ExprP c2 = expression(); expectKw("then"); s->clauses.push_back({ c2, block() }); }
            if (acceptKw("else")) s->orelse = block();
            expectKw("end"); return s;
        }
        if (acceptKw("for")) {
            std::string n1 = expectName();
            if (acceptSym("=")) {
                s->k = Stat::NumFor; s->names.push_back(n1);
                s->e1 = expression(); expectSym(","); s->e2 = expression(); if (acceptSym(",")) s->e3 = expression();
            } else {
                s->k = Stat::GenFor; s->names.push_back(n1);
                while (acceptSym(",")) s->names.push_back(expectName());
                expectKw("in"); s->exprs = exprList();
            }
            expectKw("do"); s->body = block(); expectKw("end"); return s;
        }
        if (acceptKw("while")) { s->k = Stat::While; s->e1 = expression(); expectKw("do"); s->body = block(); expectKw("end"); return s; }
        if (acceptKw("repeat")) { s->k = Stat::Repeat; s->body = block(); expectKw("until"); s->e1 = expression(); return s; }
        if (acceptKw("do")) { s->k = Stat::Do; s->body = block(); expectKw("end"); return s; }
        if (acceptKw("return")) { s->k = Stat::Return; if (!blockEnd() && !isSym(";")) s->exprs = exprList(); return s; }
        if (acceptKw("break")) { s->k = Stat::Break; return s; }
        // expression statement: call or assignment
        ExprP e = suffixed();
        if (isSym("=") || isSym(",")) {
            s->k = Stat::Assign; s->targets.push_back(e);
            while (acceptSym(",")) s->targets.push_back(suffixed());
            expectSym("="); s->exprs = exprList(); return s;
        }
        if (e->k != Expr::Call && e->k != Expr::Method) fail("syntax error near '" + tok.s + "'", tok.line);
        s->k = Stat::CallStat; s->e1 = e; return s;
    }
    std::vector<ExprP> exprList() { std::vector<ExprP> v; v.push_back(expression()); while (acceptSym(",")) v.push_back(expression()); return v; }
    ExprP primary() {
        if (tok.t == Token::Name) { ExprP e = mk(Expr::Name); e->str = tok.s; advance(); return e; }
        if (acceptSym("(")) { ExprP e = expression(); expectSym(")"); ExprP p = mk(Expr::Unop); p->str = "()"; p->a = e; return fin(p); }
        fail("unexpected symbol near '" + tok.s + "'", tok.line);
    }
    std::vector<ExprP> callArgs() {
        std::vector<ExprP> a;
        if (tok.t == Token::String) { ExprP e = mk(Expr::String); e->str = tok.s; advance(); a.push_back(e); return a; }
        if (isSym("{")) { a.push_back(tableCtor()); return a; }
        expectSym("(");
        if (!isSym(")")) a = exprList();
        expectSym(")");
        return a;
    }
    ExprP suffixed() {
        ExprP e = primary();
        for (;;) {
            if (acceptSym(".")) { ExprP i = mk(Expr::Index); i->a = e; ExprP k = mk(Expr::String); k->str = expectName(); i->b = k; e = fin(i); }
            else if (acceptSym("[")) { ExprP i = mk(Expr::Index); i->a = e; i->b = expression(); expectSym("]"); e = fin(i); }
            else if (acceptSym(":")) { ExprP m = mk(Expr::Method); m->a = e; m->str = expectName(); m->args = callArgs(); e = fin(m); }
            else if (isSym("(") || isSym("{") || tok.t == Token::String) { ExprP c = mk(Expr::Call); c->a = e; c->args = callArgs(); e = fin(c); }
            else return e;
        }
    }
    ExprP tableCtor() {
        ExprP t = mk(Expr::TableCtor);
        expectSym("{");
        while (!isSym("}")) {
            if (isSym("[")) { advance(); ExprP k = expression(); expectSym("]"); expectSym("="); t->fields.push_back({ k, expression() }); }
            else if (tok.t == Token::Name && peek().t == Token::Sym && peek().s == "=") { ExprP k = mk(Expr::String); k->str = tok.s; advance(); advance(); t->fields.push_back({ k, expression() }); }
            else t->args.push_back(expression());
            if (!acceptSym(",") && !acceptSym(";")) break;
        }
        expectSym("}");
        return fin(t);
    }
    ExprP simple() {
        if (tok.t == Token::Number) { ExprP e = mk(Expr::Number); e->num = tok.n; advance(); return e; }
        if (tok.t == Token::String) { ExprP e = mk(Expr::String); e->str = tok.s; advance(); return e; }
        if (acceptKw("nil")) return mk(Expr::Nil);
        if (acceptKw("true")) return mk(Expr::True);
        if (acceptKw("false")) return mk(Expr::False);
        if (acceptSym("...")) return mk(Expr::Vararg);
        if (isSym("{")) return tableCtor();
        if (acceptKw("function")) return functionBody(false);
        return suffixed();
    }
    static int prec(const Token& t, bool* right) {
        *right = false;
        if (t.t == Token::Kw) { if (t.s == "or") return 1; if (t.s == "and") return 2; return 0; }
        if (t.t != Token::Sym) return 0;
        const std::string& s = t.s;
        if (s == "<" || s == ">" || s == "<=" || s == ">=" || s == "~=" || s == "==") return 3;
        if (s == "..") { *right = true; return 4; }
        if (s == "+" || s == "-") return 5;
        if (s == "*" || s == "/" || s == "%") return 6;
        if (s == "^") { *right = true; return 8; }
        return 0;
    }
    ExprP expression(int limit = 0) {
        NestGuard guard(*this);
        ExprP left;
        if (isKw("not") || isSym("-") || isSym("#")) {
            ExprP u = mk(Expr::Unop); u->str = tok.s; advance(); u->a = expression(7); left = fin(u);
        } else left = simple();
        for (;;) {
            bool right; int p = prec(tok, &right);
            if (p == 0 || p <= limit) break;
            ExprP b = mk(Expr::Binop); b->str = tok.s; advance();
            b->a = left; b->b = expression(right ? p - 1 : p); left = fin(b);
        }
        return left;
    }
};

// ===================================================================================== interpreter
struct Scope { std::map<std::string, Value> vars; std::shared_ptr<Scope> parent; };
struct Closure { ExprP fn; std::shared_ptr<Scope> env; };
struct BreakEx {}; struct ReturnEx { std::vector<Value> vals; };

struct Problem {   // everything the spec declared
    Dag dag;
    std::vector<ImageDecl> images; std::vector<ParamDecl> params; std::vector<GraphDecl> graphs;
    std::vector<std::vector<int>> slotDims;      // per global graph slot
    std::vector<int> terms;                      // scalar residual expressions
    std::vector<int> excludes;
    bool usePreconditioner = false;
    int nUnknown = 0, maxDim = -1;
    int computedCounter = 0;
};

struct Interp {
    Problem& P;
    std::shared_ptr<Scope> globals = std::make_shared<Scope>();
    int line = 0;
    // resource limits: a specification is untrusted text run inside the caller's process (the reference runs it in a Lua VM)
    int depth = 0;            // nested function calls
    long steps = 0;           // statements executed
    static constexpr int kMaxDepth = 200;
    static constexpr long kMaxSteps = 5000000;
    static constexpr size_t kMaxString = 1 << 20;
    struct DepthGuard { int& d; explicit DepthGuard(int& x) : d(x) { ++d; } ~DepthGuard() { --d; } };
    explicit Interp(Problem& p) : P(p) {}

    // ---------- value helpers
    bool isExprLike(const Value& v) const { return v.type == Value::Num || (v.type == Value::Obj && (dynamic_cast<ExprObj*>(v.obj.get()) || dynamic_cast<VecObj*>(v.obj.get()))); }
    bool isVec(const Value& v) const { return v.type == Value::Obj && dynamic_cast<VecObj*>(v.obj.get()); }
    int toExpr(const Value& v, const char* what = "expression") {
        if (v.type == Value::Num) return P.dag.constant(v.n);
        if (v.type == Value::Bool) return P.dag.constant(v.b ? 1 : 0);
        if (v.type == Value::Obj) if (auto* e = dynamic_cast<ExprObj*>(v.obj.get())) return e->e;
        fail(std::string("expected a scalar ") + what, line);
    }
    std::vector<int> toVec(const Value& v) {
        if (auto* vo = v.type == Value::Obj ? dynamic_cast<VecObj*>(v.obj.get()) : nullptr) return vo->e;
        return { toExpr(v) };
    }
    Value mkExpr(int e) { double c; if (P.dag.isConst(e, &c) && false) return Value::num(c); return Value::object(std::make_shared<ExprObj>(e)); }
    Value mkVec(const std::vector<int>& e) { auto v = std::make_shared<VecObj>(); v->e = e; return Value::object(v); }
    Value fromComponents(const std::vector<int>& e) { return e.size() == 1 ? mkExpr(e[0]) : mkVec(e); }
    // element-wise lifting with scalar broadcast (ad.t:330-348)
    Value lift2(const Value& a, const Value& b, const std::function<int(int, int)>& f) {
        if (a.type == Value::Num && b.type == Value::Num) { int r = f(P.dag.constant(a.n), P.dag.constant(b.n)); double c; if (P.dag.isConst(r, &c)) return Value::num(c); return mkExpr(r); }
        std::vector<int> x = toVec(a), y = toVec(b);
        size_t n = std::max(x.size(), y.size());
        if ((x.size() != 1 && x.size() != n) || (y.size() != 1 && y.size() != n)) fail("vector size mismatch", line);
        std::vector<int> r(n);
        for (size_t i = 0; i < n; ++i) r[i] = f(x[x.size() == 1 ? 0 : i], y[y.size() == 1 ? 0 : i]);
        return (isVec(a) || isVec(b)) ? mkVec(r) : mkExpr(r[0]);
    }
    Value lift1(const Value& a, const std::function<int(int)>& f) {
        if (a.type == Value::Num) { int r = f(P.dag.constant(a.n)); double c; if (P.dag.isConst(r, &c)) return Value::num(c); return mkExpr(r); }
        std::vector<int> x = toVec(a), r(x.size());
        for (size_t i = 0; i < x.size(); ++i) r[i] = f(x[i]);
        return isVec(a) ? mkVec(r) : mkExpr(r[0]);
    }

    // ---------- scopes
    Value* lookup(const std::shared_ptr<Scope>& s, const std::string& n) {
        for (Scope* c = s.get(); c; c = c->parent.get()) { auto it = c->vars.find(n); if (it != c->vars.end()) return &it->second; }
        return nullptr;
    }
    void def(const std::string& name, std::function<std::vector<Value>(Interp&, std::vector<Value>&, int)> f) {
        Value v; v.type = Value::Builtin; v.bf = std::make_shared<std::function<std::vector<Value>(Interp&, std::vector<Value>&, int)>>(std::move(f));
        globals->vars[name] = v;
    }

    // ---------- shifting (ComputedArray access at an offset)
    std::map<std::pair<int, std::string>, int> shiftMemo;
    int shift(int e, const int off[3]) {
        if (off[0] == 0 && off[1] == 0 && off[2] == 0) return e;
        std::string key = std::to_string(off[0]) + "," + std::to_string(off[1]) + "," + std::to_string(off[2]);
        auto it = shiftMemo.find({ e, key });
        if (it != shiftMemo.end()) return it->second;
        Node n = P.dag.nodes[e];
        int r;
        switch (n.op) {
            case Op::Const: case Op::Param: r = e; break;
            case Op::Load: case Op::Unknown:
                if (n.graph) fail("a ComputedArray read through a graph element cannot be shifted", line);
                for (int k = 0; k < 3; ++k) n.off[k] += off[k];
                r = P.dag.add(n); break;
            case Op::Index: n.off[0] += off[n.id]; r = P.dag.add(n); break;
            case Op::Bounds: for (int k = 0; k < 3; ++k) { n.off[k] += off[k]; n.off2[k] += off[k]; } r = P.dag.add(n); break;
            case Op::Sample: { Node m = n; m.args = { shift(n.args[0], off), shift(n.args[1], off) }; r = P.dag.add(m); break; }
            default: {
                std::vector<int> a; for (int c : n.args) a.push_back(shift(c, off));
                if (n.op == Op::Select) r = P.dag.select(a[0], a[1], a[2]);
                else if (a.size() == 1) r = P.dag.unary(n.op, a[0]);
                else r = P.dag.binary(n.op, a[0], a[1]);
            }
        }
        shiftMemo[{ e, key }] = r;
        return r;
    }

    // ---------- calling
    std::vector<Value> call(const Value& f, std::vector<Value>& args, int ln) {
        line = ln;
        if (depth >= kMaxDepth) fail("stack overflow: function calls nested deeper than " + std::to_string(kMaxDepth), ln);
        DepthGuard guard(depth);
        if (f.type == Value::Builtin) return (*f.bf)(*this, args, ln);
        if (f.type == Value::Fn) {
            auto sc = std::make_shared<Scope>(); sc->parent = f.fn->env;
            const Expr& fn = *f.fn->fn;
            for (size_t i = 0; i < fn.params.size(); ++i) sc->vars[fn.params[i]] = i < args.size() ? args[i] : Value();
            if (fn.vararg) { auto t = std::make_shared<Table>(); for (size_t i = fn.params.size(); i < args.size(); ++i) t->arr[(double)(i - fn.params.size() + 1)] = args[i]; Value tv; tv.type = Value::Tab; tv.t = t; sc->vars["..."] = tv; }
            try { execBlock(*fn.body, sc); } catch (ReturnEx& r) { return r.vals; }
            return {};
        }
        if (f.type == Value::Obj) return callObject(f, args, ln);
        if (f.type == Value::Tab) { Value mt = f.t->get(Value::str("__call")); if (mt.type != Value::Nil) { std::vector<Value> a2; a2.push_back(f); for (auto& a : args) a2.push_back(a); return call(mt, a2, ln); } }
        fail("attempt to call a non-function value", ln);
    }
    // image(...), vector(i)
    std::vector<Value> callObject(const Value& f, std::vector<Value>& args, int ln) {
        if (auto* v = dynamic_cast<VecObj*>(f.obj.get())) {
            if (args.size() != 1 || args[0].type != Value::Num) fail("indices into vectors must be numbers", ln);
            int i = (int)args[0].n; if (i < 0 || i >= (int)v->e.size()) fail("vector index out of bounds", ln);
            return { mkExpr(v->e[i]) };
        }
        if (auto* ex = dynamic_cast<ExprObj*>(f.obj.get())) {   // scalar(0) is allowed by the reference for 1-channel images
            if (args.size() == 1 && args[0].type == Value::Num && args[0].n == 0) return { mkExpr(ex->e) };
            fail("attempt to index a scalar expression", ln);
        }
        if (auto* sm = dynamic_cast<SampledObj*>(f.obj.get())) {   // SampledImage:__call(x, y [, c])  (o.t:2467-2478)
            if (args.size() < 2) fail("sampled images take (x, y [, channel])", ln);
            int x = toExpr(args[0]), y = toExpr(args[1]);
            int chan = -1; if (args.size() > 2) { if (args[2].type != Value::Num) fail("channel must be a number", ln); chan = (int)args[2].n; }
            std::vector<int> comps;
            for (int c = 0; c < sm->channels; ++c) {
                if (chan >= 0 && c != chan) continue;
                Node n; n.op = Op::Sample; n.id = sm->image; n.off[0] = sm->dx; n.off[1] = sm->dy; n.chan = c; n.args = { x, y };
                comps.push_back(P.dag.add(n));
            }
            return { fromComponents(comps) };
        }
        auto* im = dynamic_cast<ImageObj*>(f.obj.get());
        if (!im) fail(std::string("attempt to call a ") + f.obj->kind(), ln);
        int nd = (int)im->dims.size();
        int off[3] = { 0, 0, 0 }; bool graph = false; int gslot = -1; size_t used = 0;
        if (!args.empty() && args[0].type == Value::Obj && dynamic_cast<GraphElemObj*>(args[0].obj.get())) {
            auto* ge = dynamic_cast<GraphElemObj*>(args[0].obj.get());
            if (ge->dims != im->dims) fail("graph element is in a different index space from the image", ln);
            graph = true; gslot = ge->globalSlot; used = 1;
        } else {
            if ((int)args.size() < nd) fail("image access needs one offset per dimension", ln);
            for (int k = 0; k < nd; ++k) { if (args[k].type != Value::Num) fail("image offsets must be constant numbers", ln); off[k] = (int)args[k].n; }
            used = nd;
        }
        int chan = -1;
        if (args.size() > used) { if (args[used].type != Value::Num) fail("channel must be a number", ln); chan = (int)args[used].n; if (chan < 0 || chan >= im->channels) fail("channel out of bounds", ln); }
        std::vector<int> comps;
        for (int c = 0; c < im->channels; ++c) {
            if (chan >= 0 && c != chan) continue;
            if (im->id < 0) {   // computed array: inline its definition, shifted
                if (graph) fail("ComputedArray through a graph element is not supported", ln);
                comps.push_back(shift(im->computed[c], off));
            } else {
                Node n; n.op = P.images[im->id].unknown ? Op::Unknown : Op::Load; n.id = im->id; n.graph = graph; n.slot = gslot; n.chan = c;
                for (int k = 0; k < 3; ++k) n.off[k] = off[k];
                comps.push_back(P.dag.add(n));
            }
        }
        return { fromComponents(comps) };
    }

    // ---------- expression evaluation
    Value index(const Value& o, const Value& k, int ln) {
        if (o.type == Value::Tab) return o.t->get(k);
        if (o.type == Value::Obj) {
            if (auto* v = dynamic_cast<VecObj*>(o.obj.get())) {
                if (k.type == Value::Num) { int i = (int)k.n; if (i < 0 || i >= (int)v->e.size()) fail("vector index out of bounds", ln); return mkExpr(v->e[i]); }
            }
            if (auto* g = dynamic_cast<GraphObj*>(o.obj.get())) {
                if (k.type == Value::Str) {
                    auto it = g->slotOf.find(k.s); if (it == g->slotOf.end()) fail("graph has no element '" + k.s + "'", ln);
                    auto ge = std::make_shared<GraphElemObj>(); ge->graph = g->id; ge->slot = it->second;
                    int gs = 0; for (int q = 0; q < g->id; ++q) gs += (int)P.graphs[q].slots.size();
                    ge->globalSlot = gs + it->second; ge->dims = P.graphs[g->id].slots[it->second].dims;
                    return Value::object(ge);
                }
            }
            if (k.type == Value::Str) return Value::str("__method:" + k.s);   // resolved by Method calls only
        }
        if (o.type == Value::Str && k.type == Value::Str) return Value();
        fail("attempt to index a non-table value", ln);
    }
    Value arith(const std::string& op, const Value& a, const Value& b, int ln) {
        line = ln;
        if (op == "..") {
            auto s = [&](const Value& v) { if (v.type == Value::Str) return v.s; if (v.type == Value::Num) { char buf[64]; if (v.n == (long long)v.n) snprintf(buf, sizeof buf, "%lld", (long long)v.n); else snprintf(buf, sizeof buf, "%.14g", v.n); return std::string(buf); } fail("attempt to concatenate a non-string value", ln); };
            const std::string sa = s(a), sb = s(b);
            if (sa.size() + sb.size() > kMaxString) fail("string too long", ln);
            return Value::str(sa + sb);
        }
        if (op == "==" || op == "~=") {
            bool eq = false;
            if (a.type == b.type) {
                switch (a.type) { case Value::Nil: eq = true; break; case Value::Bool: eq = a.b == b.b; break; case Value::Num: eq = a.n == b.n; break; case Value::Str: eq = a.s == b.s; break;
                    case Value::Tab: eq = a.t == b.t; break; case Value::Obj: eq = a.obj == b.obj; break; default: eq = false; }
            }
            return Value::boolean(op == "==" ? eq : !eq);
        }
        if (op == "<" || op == ">" || op == "<=" || op == ">=") {
            if (a.type == Value::Num && b.type == Value::Num) return Value::boolean(op == "<" ? a.n < b.n : op == ">" ? a.n > b.n : op == "<=" ? a.n <= b.n : a.n >= b.n);
            if (a.type == Value::Str && b.type == Value::Str) return Value::boolean(op == "<" ? a.s < b.s : op == ">" ? a.s > b.s : op == "<=" ? a.s <= b.s : a.s >= b.s);
            fail("comparison operators apply to numbers; use less()/greater()/eq() on expressions", ln);
        }
        if (a.type == Value::Num && b.type == Value::Num) {
            if (op == "+") return Value::num(a.n + b.n); if (op == "-") return Value::num(a.n - b.n); if (op == "*") return Value::num(a.n * b.n);
            if (op == "/") return Value::num(a.n / b.n); if (op == "%") return Value::num(a.n - std::floor(a.n / b.n) * b.n); if (op == "^") return Value::num(std::pow(a.n, b.n));
        }
        if (!isExprLike(a) || !isExprLike(b)) fail("arithmetic on a non-numeric value ('" + op + "')", ln);
        Op o = op == "+" ? Op::Add : op == "-" ? Op::Sub : op == "*" ? Op::Mul : op == "/" ? Op::Div : op == "^" ? Op::Pow : Op::Const;
        if (o == Op::Const) fail("operator '" + op + "' is not supported on expressions", ln);
        return lift2(a, b, [&](int x, int y) { return P.dag.binary(o, x, y); });
    }
    std::vector<Value> evalMulti(const ExprP& e, const std::shared_ptr<Scope>& sc) {
        if (e->k == Expr::Call) {
            Value f = eval(e->a, sc); std::vector<Value> args = evalArgs(e->args, sc);
            return call(f, args, e->line);
        }
        if (e->k == Expr::Method) {
            Value o = eval(e->a, sc); std::vector<Value> args = evalArgs(e->args, sc);
            return callMethod(o, e->str, args, e->line);
        }
        if (e->k == Expr::Vararg) { Value* v = lookup(sc, "..."); std::vector<Value> r; if (v && v->type == Value::Tab) for (int i = 1; i <= v->t->length(); ++i) r.push_back(v->t->arr[(double)i]); return r; }
        return { eval(e, sc) };
    }
    std::vector<Value> evalArgs(const std::vector<ExprP>& es, const std::shared_ptr<Scope>& sc) {
        std::vector<Value> r;
        for (size_t i = 0; i < es.size(); ++i) {
            if (i + 1 == es.size()) { auto m = evalMulti(es[i], sc); for (auto& v : m) r.push_back(v); }
            else r.push_back(eval(es[i], sc));
        }
        return r;
    }
    std::vector<Value> callMethod(const Value& o, const std::string& name, std::vector<Value>& args, int ln) {
        line = ln;
        if (o.type == Value::Obj) {
            if (auto* v = dynamic_cast<VecObj*>(o.obj.get())) {
                if (name == "dot") { if (args.size() != 1) fail("dot expects one argument", ln); std::vector<int> w = toVec(args[0]); if (w.size() != v->e.size()) fail("dot: size mismatch", ln);
                    int s = P.dag.constant(0); for (size_t i = 0; i < w.size(); ++i) s = P.dag.binary(Op::Add, s, P.dag.binary(Op::Mul, v->e[i], w[i])); return { mkExpr(s) }; }
                if (name == "size") return { Value::num((double)v->e.size()) };
                if (name == "sum") { int s = P.dag.constant(0); for (int c : v->e) s = P.dag.binary(Op::Add, s, c); return { mkExpr(s) }; }
                if (name == "expressions") { auto t = std::make_shared<Table>(); for (size_t i = 0; i < v->e.size(); ++i) t->arr[(double)(i + 1)] = mkExpr(v->e[i]); Value tv; tv.type = Value::Tab; tv.t = t; return { tv }; }
            }
            if (dynamic_cast<ExprObj*>(o.obj.get())) { if (name == "size") return { Value::num(1) }; }
            fail(std::string("unknown method '") + name + "' on a " + o.obj->kind(), ln);
        }
        if (o.type == Value::Tab) { Value f = o.t->get(Value::str(name)); std::vector<Value> a2; a2.push_back(o); for (auto& a : args) a2.push_back(a); return call(f, a2, ln); }
        fail("attempt to call a method on a non-object", ln);
    }
    Value eval(const ExprP& e, const std::shared_ptr<Scope>& sc) {
        line = e->line;
        switch (e->k) {
            case Expr::Nil: return Value();
            case Expr::True: return Value::boolean(true);
            case Expr::False: return Value::boolean(false);
            case Expr::Number: return Value::num(e->num);
            case Expr::String: return Value::str(e->str);
            case Expr::Name: { Value* v = lookup(sc, e->str); if (v) return *v; return Value(); }
            case Expr::Vararg: { auto m = evalMulti(e, sc); return m.empty() ? Value() : m[0]; }
            case Expr::Index: { Value o = eval(e->a, sc), k = eval(e->b, sc); return index(o, k, e->line); }
            case Expr::Call: case Expr::Method: { auto m = evalMulti(e, sc); return m.empty() ? Value() : m[0]; }
            case Expr::Function: { Value v; v.type = Value::Fn; v.fn = std::make_shared<Closure>(); v.fn->fn = e; v.fn->env = sc; return v; }
            case Expr::TableCtor: {
                auto t = std::make_shared<Table>(); int n = 0;
                for (size_t i = 0; i < e->args.size(); ++i) {
                    if (i + 1 == e->args.size()) { auto m = evalMulti(e->args[i], sc); for (auto& v : m) t->arr[(double)(++n)] = v; }
                    else t->arr[(double)(++n)] = eval(e->args[i], sc);
                }
                for (auto& f : e->fields) t->set(eval(f.first, sc), eval(f.second, sc));
                Value v; v.type = Value::Tab; v.t = t; return v;
            }
            case Expr::Unop: {
                if (e->str == "()") return eval(e->a, sc);
                Value a = eval(e->a, sc);
                if (e->str == "not") return Value::boolean(!a.truthy());
                if (e->str == "#") { if (a.type == Value::Tab) return Value::num(a.t->length()); if (a.type == Value::Str) return Value::num((double)a.s.size()); fail("attempt to get the length of a non-table", e->line); }
                if (a.type == Value::Num) return Value::num(-a.n);
                return lift1(a, [&](int x) { return P.dag.unary(Op::Neg, x); });
            }
            case Expr::Binop: {
                if (e->str == "and") { Value a = eval(e->a, sc); return a.truthy() ? eval(e->b, sc) : a; }
                if (e->str == "or") { Value a = eval(e->a, sc); return a.truthy() ? a : eval(e->b, sc); }
                Value a = eval(e->a, sc), b = eval(e->b, sc);
                return arith(e->str, a, b, e->line);
            }
        }
        return Value();
    }
    void assign(const ExprP& target, const Value& v, const std::shared_ptr<Scope>& sc) {
        if (target->k == Expr::Name) { Value* p = lookup(sc, target->str); if (p) *p = v; else globals->vars[target->str] = v; return; }
        if (target->k == Expr::Index) { Value o = eval(target->a, sc), k = eval(target->b, sc); if (o.type != Value::Tab) fail("attempt to assign a field of a non-table", target->line); o.t->set(k, v); return; }
        fail("cannot assign to this expression", target->line);
    }
    void execBlock(const Block& b, const std::shared_ptr<Scope>& sc) { for (const StatP& s : b) exec(s, sc); }
    void exec(const StatP& s, const std::shared_ptr<Scope>& sc) {
        line = s->line;
        if (++steps > kMaxSteps) fail("the specification did not finish: more than " + std::to_string(kMaxSteps) + " statements executed", s->line);
        switch (s->k) {
            case Stat::Local: {
                std::vector<Value> vals = evalArgs(s->exprs, sc);
                for (size_t i = 0; i < s->names.size(); ++i) sc->vars[s->names[i]] = i < vals.size() ? vals[i] : Value();
                break;
            }
            case Stat::Assign: {
                std::vector<Value> vals = evalArgs(s->exprs, sc);
                for (size_t i = 0; i < s->targets.size(); ++i) assign(s->targets[i], i < vals.size() ? vals[i] : Value(), sc);
                break;
            }
            case Stat::CallStat: evalMulti(s->e1, sc); break;
            case Stat::Do: {
                if (!s->names.empty() && s->names[0] == "__splice__") { execBlock(*s->body, sc); break; }   // local function
                auto inner = std::make_shared<Scope>(); inner->parent = sc; execBlock(*s->body, inner); break;
            }
            case Stat::While: { int guard = 0; while (eval(s->e1, sc).truthy()) { auto inner = std::make_shared<Scope>(); inner->parent = sc; try { execBlock(*s->body, inner); } catch (BreakEx&) { break; } if (++guard > 10000000) fail("loop limit", s->line); } break; }
            case Stat::Repeat: { int guard = 0; for (;;) { auto inner = std::make_shared<Scope>(); inner->parent = sc; try { execBlock(*s->body, inner); } catch (BreakEx&) { break; } if (eval(s->e1, inner).truthy()) break; if (++guard > 10000000) fail("loop limit", s->line); } break; }
            case Stat::If: {
                bool done = false;
                for (auto& c : s->clauses) if (eval(c.first, sc).truthy()) { auto inner = std::make_shared<Scope>(); inner->parent = sc; execBlock(*c.second, inner); done = true; break; }
                if (!done && s->orelse) { auto inner = std::make_shared<Scope>(); inner->parent = sc; execBlock(*s->orelse, inner); }
                break;
            }
            case Stat::NumFor: {
                Value a = eval(s->e1, sc), b = eval(s->e2, sc), st = s->e3 ? eval(s->e3, sc) : Value::num(1);
                if (a.type != Value::Num || b.type != Value::Num || st.type != Value::Num || st.n == 0) fail("'for' bounds must be numbers", s->line);
                for (double i = a.n; st.n > 0 ? i <= b.n : i >= b.n; i += st.n) {
                    auto inner = std::make_shared<Scope>(); inner->parent = sc; inner->vars[s->names[0]] = Value::num(i);
                    try { execBlock(*s->body, inner); } catch (BreakEx&) { break; }
                }
                break;
            }
            case Stat::GenFor: {
                std::vector<Value> it = evalArgs(s->exprs, sc);
                if (it.empty()) fail("'for in' needs an iterator", s->line);
                Value f = it[0], state = it.size() > 1 ? it[1] : Value(), ctl = it.size() > 2 ? it[2] : Value();
                for (int guard = 0; guard < 10000000; ++guard) {
                    std::vector<Value> args = { state, ctl };
                    std::vector<Value> r = call(f, args, s->line);
                    if (r.empty() || r[0].type == Value::Nil) break;
                    ctl = r[0];
                    auto inner = std::make_shared<Scope>(); inner->parent = sc;
                    for (size_t i = 0; i < s->names.size(); ++i) inner->vars[s->names[i]] = i < r.size() ? r[i] : Value();
                    try { execBlock(*s->body, inner); } catch (BreakEx&) { break; }
                }
                break;
            }
            case Stat::Return: { ReturnEx r; r.vals = evalArgs(s->exprs, sc); throw r; }
            case Stat::Break: throw BreakEx();
        }
    }
};

// ===================================================================================== DSL prelude
std::vector<int> dimsFromTable(Interp& I, const Value& v, int ln) {
    if (v.type != Value::Tab) fail("expected a table of dimensions {W,H,...}", ln);
    std::vector<int> d;
    for (int i = 1; i <= v.t->length(); ++i) {
        Value e = v.t->arr[(double)i];
        auto* dm = e.type == Value::Obj ? dynamic_cast<DimObj*>(e.obj.get()) : nullptr;
        if (!dm) fail("dimension tables must contain Dim(...) objects", ln);
        d.push_back(dm->idx);
    }
    if (d.empty() || d.size() > 3) fail("index spaces must have 1 to 3 dimensions", ln);
    (void)I;
    return d;
}

void installPrelude(Interp& I) {
    Problem& P = I.P;
    auto num = [](const Value& v, int ln, const char* what) { if (v.type != Value::Num) fail(std::string("expected a number for ") + what, ln); return v.n; };
    auto str = [](const Value& v, int ln, const char* what) { if (v.type != Value::Str) fail(std::string("expected a string for ") + what, ln); return v.s; };
    // ---- scalar / vector types
    auto deftype = [&](const std::string& base, ScalarType st) {
        I.globals->vars[base] = Value::object(std::make_shared<TypeObj>(st, 1));
        for (int c = 2; c <= 9; ++c) I.globals->vars[base + std::to_string(c)] = Value::object(std::make_shared<TypeObj>(st, c));
    };
    deftype("opt_float", ScalarType::OptFloat); deftype("float", ScalarType::Float); deftype("double", ScalarType::Double);
    I.globals->vars["uint8"] = Value::object(std::make_shared<TypeObj>(ScalarType::UInt8, 1));
    I.globals->vars["int"] = Value::object(std::make_shared<TypeObj>(ScalarType::Int32, 1));
    I.globals->vars["int32"] = I.globals->vars["int"];

    I.def("Dim", [&P, num, str](Interp&, std::vector<Value>& a, int ln) {
        if (a.size() < 2) fail("Dim(name, index)", ln);
        auto d = std::make_shared<DimObj>(); d->name = str(a[0], ln, "Dim name"); d->idx = (int)num(a[1], ln, "Dim index");
        if (d->idx < 0 || d->idx >= 8) fail("Dim index out of range", ln);
        P.maxDim = std::max(P.maxDim, d->idx);
        return std::vector<Value>{ Value::object(d) };
    });
    auto declImage = [&P, num, str](Interp& I2, std::vector<Value>& a, int ln, bool unknown) {
        if (a.size() < 4) fail("Unknown/Array(name, type, {dims}, index)", ln);
        auto* ty = a[1].type == Value::Obj ? dynamic_cast<TypeObj*>(a[1].obj.get()) : nullptr;
        if (!ty) fail("expected a type (opt_float, opt_float3, float, uint8, ...)", ln);
        ImageDecl d; d.name = str(a[0], ln, "image name"); d.type = ty->st; d.channels = ty->channels; d.dims = dimsFromTable(I2, a[2], ln);
        d.paramIndex = (int)num(a[3], ln, "parameter index"); d.unknown = unknown;
        if (unknown) { if (d.type != ScalarType::OptFloat && d.type != ScalarType::Float) fail("unknowns must be opt_float vectors", ln); d.unknownId = P.nUnknown++; }
        P.images.push_back(d);
        auto im = std::make_shared<ImageObj>(); im->id = (int)P.images.size() - 1; im->dims = d.dims; im->channels = d.channels;
        return std::vector<Value>{ Value::object(im) };
    };
    I.def("Unknown", [declImage](Interp& I2, std::vector<Value>& a, int ln) { return declImage(I2, a, ln, true); });
    I.def("Array", [declImage](Interp& I2, std::vector<Value>& a, int ln) { return declImage(I2, a, ln, false); });
    I.globals->vars["Image"] = I.globals->vars["Array"];
    I.def("Param", [&P, num, str](Interp& I2, std::vector<Value>& a, int ln) {
        if (a.size() < 3) fail("Param(name, type, index)", ln);
        auto* ty = a[1].type == Value::Obj ? dynamic_cast<TypeObj*>(a[1].obj.get()) : nullptr;
        if (!ty || ty->channels != 1) fail("Param type must be a scalar type", ln);
        ParamDecl d; d.name = str(a[0], ln, "Param name"); d.type = ty->st; d.paramIndex = (int)num(a[2], ln, "parameter index");
        P.params.push_back(d);
        Node n; n.op = Op::Param; n.id = (int)P.params.size() - 1;
        return std::vector<Value>{ I2.mkExpr(P.dag.add(n)) };
    });
    I.def("Graph", [&P, num, str](Interp& I2, std::vector<Value>& a, int ln) {
        if (a.size() < 5 || (a.size() - 2) % 3 != 0) fail("Graph(name, countIndex, \"v0\", {dims}, index0, ...)", ln);
        GraphDecl g; g.name = str(a[0], ln, "Graph name"); g.countIndex = (int)num(a[1], ln, "edge-count index");
        auto go = std::make_shared<GraphObj>(); go->id = (int)P.graphs.size();
        for (size_t i = 2; i + 2 < a.size() + 0 && i + 2 <= a.size() - 1 + 0; i += 3) {
            GraphDecl::Slot s; s.name = str(a[i], ln, "graph element name"); s.dims = dimsFromTable(I2, a[i + 1], ln); s.paramIndex = (int)num(a[i + 2], ln, "graph element index");
            go->slotOf[s.name] = (int)g.slots.size(); g.slots.push_back(s); P.slotDims.push_back(s.dims);
        }
        P.graphs.push_back(g);
        return std::vector<Value>{ Value::object(go) };
    });
    I.def("ComputedArray", [&P, str](Interp& I2, std::vector<Value>& a, int ln) {
        if (a.size() < 3) fail("ComputedArray(name, {dims}, expression)", ln);
        (void)str(a[0], ln, "ComputedArray name");
        auto im = std::make_shared<ImageObj>(); im->id = -1; im->dims = dimsFromTable(I2, a[1], ln);
        im->computed = I2.toVec(a[2]); im->channels = (int)im->computed.size();
        ++P.computedCounter;
        return std::vector<Value>{ Value::object(im) };
    });
    I.globals->vars["ComputedImage"] = I.globals->vars["ComputedArray"];
    I.def("Energy", [&P](Interp& I2, std::vector<Value>& a, int ln) {
        for (auto& v : a) { if (!I2.isExprLike(v)) fail("Energy() expects expressions", ln); for (int e : I2.toVec(v)) P.terms.push_back(e); }
        return std::vector<Value>{};
    });
    I.def("Exclude", [&P](Interp& I2, std::vector<Value>& a, int ln) { if (a.size() != 1) fail("Exclude(expression)", ln); P.excludes.push_back(I2.toExpr(a[0])); return std::vector<Value>{}; });
    I.def("UsePreconditioner", [&P](Interp&, std::vector<Value>& a, int) { P.usePreconditioner = !a.empty() && a[0].truthy(); return std::vector<Value>{}; });
    I.def("Stencil", [](Interp&, std::vector<Value>& a, int ln) {
        if (a.size() != 1 || a[0].type != Value::Tab) fail("Stencil { {dx,dy}, ... }", ln);
        auto lst = a[0].t; auto pos = std::make_shared<int>(0);
        Value it; it.type = Value::Builtin;
        it.bf = std::make_shared<std::function<std::vector<Value>(Interp&, std::vector<Value>&, int)>>([lst, pos](Interp&, std::vector<Value>&, int) {
            ++*pos; auto f = lst->arr.find((double)*pos);
            if (f == lst->arr.end() || f->second.type != Value::Tab) return std::vector<Value>{ Value() };
            std::vector<Value> r; for (int i = 1; i <= f->second.t->length(); ++i) r.push_back(f->second.t->arr[(double)i]);
            return r;
        });
        return std::vector<Value>{ it };
    });
    auto bounds = [&P, num](Interp& I2, std::vector<Value>& a, int ln, bool expanded) {
        int nd = (int)a.size() - (expanded ? 1 : 0);
        if (nd < 1 || nd > 3) fail("InBounds expects one offset per dimension", ln);
        int ex = expanded ? (int)num(a.back(), ln, "expansion") : 0;
        Node n; n.op = Op::Bounds; n.nd = nd;
        for (int k = 0; k < nd; ++k) { int o = (int)num(a[k], ln, "offset"); n.off[k] = o - ex; n.off2[k] = o + ex; }
        return std::vector<Value>{ I2.mkExpr(P.dag.add(n)) };
    };
    I.def("InBounds", [bounds](Interp& I2, std::vector<Value>& a, int ln) { return bounds(I2, a, ln, false); });
    I.def("InBoundsExpanded", [bounds](Interp& I2, std::vector<Value>& a, int ln) { return bounds(I2, a, ln, true); });
    I.def("Index", [&P, num](Interp& I2, std::vector<Value>& a, int ln) { Node n; n.op = Op::Index; n.id = (int)num(a.at(0), ln, "dimension"); return std::vector<Value>{ I2.mkExpr(P.dag.add(n)) }; });
    I.def("Select", [&P](Interp& I2, std::vector<Value>& a, int ln) {
        if (a.size() != 3) fail("Select(condition, a, b)", ln);
        std::vector<int> c = I2.toVec(a[0]), x = I2.toVec(a[1]), y = I2.toVec(a[2]);
        size_t n = std::max(c.size(), std::max(x.size(), y.size()));
        auto at = [&](const std::vector<int>& v, size_t i) { if (v.size() != 1 && v.size() != n) fail("Select: vector size mismatch", ln); return v[v.size() == 1 ? 0 : i]; };
        std::vector<int> r(n); for (size_t i = 0; i < n; ++i) r[i] = P.dag.select(at(c, i), at(x, i), at(y, i));
        return std::vector<Value>{ n == 1 && !I2.isVec(a[0]) && !I2.isVec(a[1]) && !I2.isVec(a[2]) ? I2.mkExpr(r[0]) : I2.mkVec(r) };
    });
    I.globals->vars["select_"] = I.globals->vars["Select"];
    auto cmp = [&](const char* name, Op op) { I.def(name, [&P, op, name](Interp& I2, std::vector<Value>& a, int ln) { if (a.size() != 2) fail(std::string(name) + "(a, b)", ln); return std::vector<Value>{ I2.lift2(a[0], a[1], [&](int x, int y) { return P.dag.binary(op, x, y); }) }; }); };
    cmp("less", Op::Less); cmp("lesseq", Op::LessEq); cmp("greater", Op::Greater); cmp("greatereq", Op::GreaterEq); cmp("eq", Op::Eq);
    auto un = [&](const char* name, Op op) { I.def(name, [&P, op, name](Interp& I2, std::vector<Value>& a, int ln) { if (a.size() != 1) fail(std::string(name) + "(x)", ln); return std::vector<Value>{ I2.lift1(a[0], [&](int x) { return P.dag.unary(op, x); }) }; }); };
    un("sqrt", Op::Sqrt); un("Sqrt", Op::Sqrt); un("sin", Op::Sin); un("cos", Op::Cos); un("abs", Op::Abs); un("exp", Op::Exp); un("log", Op::Log); un("Not", Op::Not); un("not_", Op::Not);
    for (const Math1& m : kMath1) un(m.name, m.op);
    I.def("pow", [&P](Interp& I2, std::vector<Value>& a, int ln) { if (a.size() != 2) fail("pow(a, b)", ln); return std::vector<Value>{ I2.lift2(a[0], a[1], [&](int x, int y) { return P.dag.binary(Op::Pow, x, y); }) }; });
    I.def("And", [&P](Interp& I2, std::vector<Value>& a, int) { int r = P.dag.constant(1); for (auto& v : a) r = P.dag.binary(Op::Mul, r, I2.toExpr(v)); return std::vector<Value>{ I2.mkExpr(r) }; });
    I.globals->vars["and_"] = I.globals->vars["And"];
    I.def("Or", [&P](Interp& I2, std::vector<Value>& a, int) { int r = P.dag.constant(0); for (auto& v : a) r = P.dag.binary(Op::Or, r, I2.toExpr(v)); return std::vector<Value>{ I2.mkExpr(r) }; });
    I.globals->vars["or_"] = I.globals->vars["Or"];
    I.def("All", [&P](Interp& I2, std::vector<Value>& a, int ln) { if (a.size() != 1) fail("All(vector)", ln); int r = P.dag.constant(1); for (int e : I2.toVec(a[0])) r = P.dag.binary(Op::Mul, r, e); return std::vector<Value>{ I2.mkExpr(r) }; });
    I.def("Vector", [](Interp& I2, std::vector<Value>& a, int) { std::vector<int> e; for (auto& v : a) for (int c : I2.toVec(v)) e.push_back(c); return std::vector<Value>{ I2.mkVec(e) }; });
    auto vec3 = [](Interp& I2, const Value& v, int ln, size_t n) { std::vector<int> e = I2.toVec(v); if (e.size() != n) fail("expected a " + std::to_string(n) + "-vector", ln); return e; };
    auto add = [&P](int a, int b) { return P.dag.binary(Op::Add, a, b); }; auto sub = [&P](int a, int b) { return P.dag.binary(Op::Sub, a, b); };
    auto mul = [&P](int a, int b) { return P.dag.binary(Op::Mul, a, b); }; auto neg = [&P](int a) { return P.dag.unary(Op::Neg, a); };
    I.def("Dot3", [vec3, add, mul](Interp& I2, std::vector<Value>& a, int ln) { auto x = vec3(I2, a.at(0), ln, 3), y = vec3(I2, a.at(1), ln, 3); return std::vector<Value>{ I2.mkExpr(add(add(mul(x[0], y[0]), mul(x[1], y[1])), mul(x[2], y[2]))) }; });
    I.def("Matrix3x3Mul", [vec3, add, mul](Interp& I2, std::vector<Value>& a, int ln) {
        auto m = vec3(I2, a.at(0), ln, 9), v = vec3(I2, a.at(1), ln, 3);
        std::vector<int> r; for (int i = 0; i < 3; ++i) r.push_back(add(add(mul(m[3 * i], v[0]), mul(m[3 * i + 1], v[1])), mul(m[3 * i + 2], v[2])));
        return std::vector<Value>{ I2.mkVec(r) };
    });
    I.def("normalize", [&P, vec3, add, mul](Interp& I2, std::vector<Value>& a, int ln) {
        auto v = vec3(I2, a.at(0), ln, 3); int l = P.dag.unary(Op::Sqrt, add(add(mul(v[0], v[0]), mul(v[1], v[1])), mul(v[2], v[2])));
        return std::vector<Value>{ I2.mkVec({ P.dag.binary(Op::Div, v[0], l), P.dag.binary(Op::Div, v[1], l), P.dag.binary(Op::Div, v[2], l) }) };
    });
    I.def("length", [&P, vec3, add, mul, sub](Interp& I2, std::vector<Value>& a, int ln) {
        auto x = vec3(I2, a.at(0), ln, 3), y = vec3(I2, a.at(1), ln, 3); int d0 = sub(x[0], y[0]), d1 = sub(x[1], y[1]), d2 = sub(x[2], y[2]);
        return std::vector<Value>{ I2.mkExpr(P.dag.unary(Op::Sqrt, add(add(mul(d0, d0), mul(d1, d1)), mul(d2, d2)))) };
    });
    I.def("L_2_norm", [&P, add, mul](Interp& I2, std::vector<Value>& a, int) { int s = P.dag.constant(0); for (int e : I2.toVec(a.at(0))) s = add(s, mul(e, e)); return std::vector<Value>{ I2.mkExpr(P.dag.unary(Op::Sqrt, s)) }; });
    I.def("Rotate2D", [&P, vec3, add, sub, mul](Interp& I2, std::vector<Value>& a, int ln) {   // lib.t:92-95
        int ang = I2.toExpr(a.at(0)); auto v = vec3(I2, a.at(1), ln, 2); int c = P.dag.unary(Op::Cos, ang), s = P.dag.unary(Op::Sin, ang);
        return std::vector<Value>{ I2.mkVec({ sub(mul(c, v[0]), mul(s, v[1])), add(mul(s, v[0]), mul(c, v[1])) }) };
    });
    I.def("Rotate3D", [&P, vec3, add, sub, mul, neg](Interp& I2, std::vector<Value>& a, int ln) {   // lib.t:77-91
        auto an = vec3(I2, a.at(0), ln, 3), v = vec3(I2, a.at(1), ln, 3);
        int ca = P.dag.unary(Op::Cos, an[0]), cb = P.dag.unary(Op::Cos, an[1]), cg = P.dag.unary(Op::Cos, an[2]);
        int sa = P.dag.unary(Op::Sin, an[0]), sb = P.dag.unary(Op::Sin, an[1]), sg = P.dag.unary(Op::Sin, an[2]);
        int m[9] = { mul(cg, cb), add(mul(neg(sg), ca), mul(mul(cg, sb), sa)), add(mul(sg, sa), mul(mul(cg, sb), ca)),
                     mul(sg, cb), add(mul(cg, ca), mul(mul(sg, sb), sa)), add(mul(neg(cg), sa), mul(mul(sg, sb), ca)),
                     neg(sb), mul(cb, sa), mul(cb, ca) };
        std::vector<int> r; for (int i = 0; i < 3; ++i) r.push_back(add(add(mul(m[3 * i], v[0]), mul(m[3 * i + 1], v[1])), mul(m[3 * i + 2], v[2])));
        return std::vector<Value>{ I2.mkVec(r) };
    });
    I.def("SampledImage", [](Interp&, std::vector<Value>& a, int ln) {   // o.t:2486-2501
        auto img = [&](const Value& v) -> ImageObj* { auto* im = v.type == Value::Obj ? dynamic_cast<ImageObj*>(v.obj.get()) : nullptr; if (!im || im->id < 0 || im->dims.size() != 2) fail("sampled images must be 2-D Arrays", ln); return im; };
        if (a.empty()) fail("SampledImage(image [, imagedx, imagedy])", ln);
        auto s = std::make_shared<SampledObj>(); ImageObj* m = img(a[0]); s->image = m->id; s->channels = m->channels;
        if (a.size() >= 3) { s->dx = img(a[1])->id; s->dy = img(a[2])->id; }
        return std::vector<Value>{ Value::object(s) };
    });
    // lib.t:105-115: L_p(val, val_const, p, dims) = ComputedArray( sqrt( (|val_const| + eps)^(p-2) ) )(0,0) * val  -- the weight is a function of
    // constant data only, so inlining it is exact
    I.def("L_p", [&P](Interp& I2, std::vector<Value>& a, int ln) {
        if (a.size() < 3) fail("L_p(val, val_const, p, dims)", ln);
        int s = P.dag.constant(0); for (int e : I2.toVec(a[1])) s = P.dag.binary(Op::Add, s, P.dag.binary(Op::Mul, e, e));
        int dist = P.dag.unary(Op::Sqrt, s);
        int C = P.dag.binary(Op::Pow, P.dag.binary(Op::Add, dist, P.dag.constant(0.0000001)), P.dag.binary(Op::Sub, I2.toExpr(a[2]), P.dag.constant(2)));
        int w = P.dag.unary(Op::Sqrt, C);
        return std::vector<Value>{ I2.lift1(a[0], [&](int v) { return P.dag.binary(Op::Mul, w, v); }) };
    });
    // lib.t:60-74: Slice(image, s, e) is callable like the image and yields channels [s, e) (a scalar when e == s + 1)
    I.def("Slice", [num](Interp&, std::vector<Value>& a, int ln) {
        if (a.size() != 3) fail("Slice(image, firstChannel, endChannel)", ln);
        const Value im = a[0]; const int s = (int)num(a[1], ln, "first channel"), e = (int)num(a[2], ln, "end channel");
        if (s < 0 || e <= s) fail("Slice: empty or negative channel range", ln);
        Value f; f.type = Value::Builtin;
        f.bf = std::make_shared<std::function<std::vector<Value>(Interp&, std::vector<Value>&, int)>>([im, s, e](Interp& I3, std::vector<Value>& args, int ln2) {
            std::vector<Value> r = I3.call(im, args, ln2);
            if (r.empty()) fail("Slice: the image access returned nothing", ln2);
            std::vector<int> v = I3.toVec(r[0]);
            if ((int)v.size() < e) fail("Slice: channel range exceeds the image's channels", ln2);
            if (s + 1 == e) return std::vector<Value>{ I3.mkExpr(v[s]) };
            return std::vector<Value>{ I3.mkVec(std::vector<int>(v.begin() + s, v.begin() + e)) };
        });
        return std::vector<Value>{ f };
    });
    // ---- a little of the Lua standard library
    I.def("print", [](Interp&, std::vector<Value>&, int) { return std::vector<Value>{}; });
    I.def("tostring", [](Interp&, std::vector<Value>& a, int) { Value v = a.empty() ? Value() : a[0]; if (v.type == Value::Num) { char b[64]; if (v.n == (long long)v.n) snprintf(b, sizeof b, "%lld", (long long)v.n); else snprintf(b, sizeof b, "%.14g", v.n); return std::vector<Value>{ Value::str(b) }; } if (v.type == Value::Str) return std::vector<Value>{ v }; return std::vector<Value>{ Value::str("<value>") }; });
    I.def("tonumber", [](Interp&, std::vector<Value>& a, int) { Value v = a.empty() ? Value() : a[0]; if (v.type == Value::Num) return std::vector<Value>{ v }; if (v.type == Value::Str) { char* e; double d = strtod(v.s.c_str(), &e); if (e != v.s.c_str()) return std::vector<Value>{ Value::num(d) }; } return std::vector<Value>{ Value() }; });
    I.def("type", [](Interp&, std::vector<Value>& a, int) { static const char* n[] = { "nil", "boolean", "number", "string", "table", "function", "function", "userdata" }; return std::vector<Value>{ Value::str(n[a.empty() ? 0 : (int)a[0].type]) }; });
    I.def("assert", [](Interp&, std::vector<Value>& a, int ln) { if (a.empty() || !a[0].truthy()) fail(a.size() > 1 && a[1].type == Value::Str ? a[1].s : "assertion failed", ln); return a; });
    I.def("error", [](Interp&, std::vector<Value>& a, int ln) -> std::vector<Value> { fail(!a.empty() && a[0].type == Value::Str ? a[0].s : "error", ln); });
    I.def("ipairs", [](Interp&, std::vector<Value>& a, int ln) {
        if (a.empty() || a[0].type != Value::Tab) fail("ipairs expects a table", ln);
        auto t = a[0].t; Value it; it.type = Value::Builtin;
        it.bf = std::make_shared<std::function<std::vector<Value>(Interp&, std::vector<Value>&, int)>>([t](Interp&, std::vector<Value>& s, int) {
            double i = (s.size() > 1 && s[1].type == Value::Num ? s[1].n : 0) + 1; auto f = t->arr.find(i);
            if (f == t->arr.end()) return std::vector<Value>{ Value() };
            return std::vector<Value>{ Value::num(i), f->second };
        });
        return std::vector<Value>{ it, a[0], Value::num(0) };
    });
    I.def("unpack", [](Interp&, std::vector<Value>& a, int ln) { if (a.empty() || a[0].type != Value::Tab) fail("unpack expects a table", ln); std::vector<Value> r; for (int i = 1; i <= a[0].t->length(); ++i) r.push_back(a[0].t->arr[(double)i]); return r; });
    {
        auto m = std::make_shared<Table>();
        auto mf = [&](const char* n, double (*f)(double)) { Value v; v.type = Value::Builtin; v.bf = std::make_shared<std::function<std::vector<Value>(Interp&, std::vector<Value>&, int)>>([f, n](Interp&, std::vector<Value>& a, int ln) { if (a.empty() || a[0].type != Value::Num) fail(std::string("math.") + n + " expects a number", ln); return std::vector<Value>{ Value::num(f(a[0].n)) }; }); m->fields[n] = v; };
        mf("sqrt", std::sqrt); mf("sin", std::sin); mf("cos", std::cos); mf("floor", std::floor); mf("ceil", std::ceil); mf("abs", std::fabs); mf("exp", std::exp); mf("log", std::log);
        m->fields["pi"] = Value::num(3.14159265358979323846); m->fields["huge"] = Value::num(HUGE_VAL);
        Value mv; mv.type = Value::Tab; mv.t = m; I.globals->vars["math"] = mv;
        auto tb = std::make_shared<Table>();
        Value ins; ins.type = Value::Builtin; ins.bf = std::make_shared<std::function<std::vector<Value>(Interp&, std::vector<Value>&, int)>>([](Interp&, std::vector<Value>& a, int ln) { if (a.size() < 2 || a[0].type != Value::Tab) fail("table.insert(t, v)", ln); a[0].t->arr[(double)(a[0].t->length() + 1)] = a.back(); return std::vector<Value>{}; });
        tb->fields["insert"] = ins; tb->fields["unpack"] = I.globals->vars["unpack"];
        Value tv; tv.type = Value::Tab; tv.t = tb; I.globals->vars["table"] = tv;
    }
    // `opt.Dim(...)` etc. (some specs qualify the DSL with opt.)
    {
        auto o = std::make_shared<Table>();
        for (const char* n : { "Dim", "Unknown", "Array", "Image", "Param", "Graph", "ComputedArray", "InBounds", "InBoundsExpanded", "Index" }) o->fields[n] = I.globals->vars[n];
        Value ov; ov.type = Value::Tab; ov.t = o; I.globals->vars["opt"] = ov;
    }
}

// ===================================================================================== analysis + code generation
struct TermInfo { bool centered = true; std::vector<int> dims; int graph = -1; std::set<std::string> deps; bool usesBounds = false; int bmin[3] = { 0, 0, 0 }, bmax[3] = { 0, 0, 0 }; bool any = false; };

void analyse(const Problem& P, int e, TermInfo& t, std::set<int>& seen) {
    if (!seen.insert(e).second) return;
    const Node& n = P.dag.nodes[e];
    if (n.op == Op::Load || n.op == Op::Unknown) {
        const ImageDecl& im = P.images[n.id];
        if (n.graph) {
            int g = 0, s = n.slot; while (s >= (int)P.graphs[g].slots.size()) { s -= (int)P.graphs[g].slots.size(); ++g; }
            if (t.any && (t.centered || t.graph != g)) fail("residual contains image reads from multiple domains");
            t.centered = false; t.graph = g;
        } else {
            if (t.any && (!t.centered || t.dims != im.dims)) fail("residual contains image reads from multiple domains");
            t.centered = true; t.dims = im.dims;
            for (int k = 0; k < 3; ++k) { t.bmin[k] = std::min(t.bmin[k], n.off[k]); t.bmax[k] = std::max(t.bmax[k], n.off[k]); }
        }
        t.any = true;
        if (n.op == Op::Unknown) t.deps.insert(std::to_string(n.id) + ":" + std::to_string(n.graph ? 1000 + n.slot : 0) + ":" + std::to_string(n.off[0]) + "," + std::to_string(n.off[1]) + "," + std::to_string(n.off[2]) + ":" + std::to_string(n.chan));
    }
    if (n.op == Op::Bounds) t.usesBounds = true;
    for (int a : n.args) analyse(P, a, t, seen);
}

struct Codegen {
    const Problem& P;
    std::ostringstream out;
    std::vector<char> dual;    // per node: depends on an unknown
    explicit Codegen(const Problem& p) : P(p), dual(p.dag.nodes.size(), 0) {
        for (size_t i = 0; i < P.dag.nodes.size(); ++i) {   // nodes are created after their arguments => one forward pass
            const Node& n = P.dag.nodes[i]; char d = n.op == Op::Unknown;
            for (int a : n.args) d = d || dual[a];
            dual[i] = d;
        }
    }
    static std::string lit(double v) {
        if (std::isinf(v)) return v > 0 ? "T(1e300*1e300)" : "T(-1e300*1e300)";
        if (std::isnan(v)) return "T((1e300*1e300) * 0.0)";   // a folded domain error (asin(2), log(-1)...) stays a NaN, as it would be at run time
        char b[64]; snprintf(b, sizeof b, "%.17g", v); std::string s = b;
        if (s.find_first_of(".en") == std::string::npos) s += ".0";
        return "T(" + s + ")";
    }
    static const char* ctype(ScalarType t) { switch (t) { case ScalarType::OptFloat: return "T"; case ScalarType::Float: return "float"; case ScalarType::Double: return "double"; case ScalarType::UInt8: return "unsigned char"; default: return "int"; } }
    std::string v(int e) const { return "v" + std::to_string(e); }
    // emits the statements computing node e (and everything below it) once; `valueMode`: unknowns are plain loads
    void emit(int e, std::set<int>& done, const TermInfo& t, bool valueMode) {
        if (!done.insert(e).second) return;
        const Node& n = P.dag.nodes[e];
        for (int a : n.args) emit(a, done, t, valueMode);
        const bool d = dual[e] && !valueMode;
        const std::string ty = d ? "S" : "T";
        out << "            const " << ty << " " << v(e) << " = ";
        auto A = [&](int k) { return v(n.args[k]); };
        auto val = [&](int k) { return (dual[n.args[k]] && !valueMode) ? "geVal(" + v(n.args[k]) + ")" : v(n.args[k]); };
        switch (n.op) {
            case Op::Const: out << lit(n.value); break;
            case Op::Param: out << "(T)P.par[" << n.id << "]"; break;
            case Op::Index: out << "T(i" << n.id << (n.off[0] >= 0 ? " + " : " - ") << std::abs(n.off[0]) << ")"; break;
            case Op::Bounds: {
                out << "((";
                for (int k = 0; k < n.nd; ++k) out << (k ? " && " : "") << "i" << k << " + (" << n.off[k] << ") >= 0 && i" << k << " + (" << n.off2[k] << ") < n" << k;
                out << ") ? T(1) : T(0))"; break;
            }
            case Op::Load: case Op::Unknown: {
                const ImageDecl& im = P.images[n.id];
                std::string idx, inb;
                if (n.graph) { idx = "e" + std::to_string(n.slot); inb = "true"; }
                else {
                    int nd = (int)im.dims.size();
                    for (int k = 0; k < nd; ++k) inb += std::string(k ? " && " : "") + "i" + std::to_string(k) + " + (" + std::to_string(n.off[k]) + ") >= 0 && i" + std::to_string(k) + " + (" + std::to_string(n.off[k]) + ") < n" + std::to_string(k);
                    idx = "(i0 + (" + std::to_string(n.off[0]) + "))";
                    if (nd > 1) idx += " + n0 * ((i1 + (" + std::to_string(n.off[1]) + "))" + (nd > 2 ? " + n1 * (i2 + (" + std::to_string(n.off[2]) + "))" : "") + ")";
                }
                if (n.op == Op::Unknown && !valueMode) out << "((" << inb << ") ? X(" << im.unknownId << ", " << idx << ", " << n.chan << ") : S(T(0)))";
                else out << "((" << inb << ") ? (T)((const " << ctype(im.type) << "*)P.img[" << n.id << "])[(" << idx << ") * " << im.channels << " + " << n.chan << "] : T(0))";
                break;
            }
            case Op::Add: out << A(0) << " + " << A(1); break;
            case Op::Sub: out << A(0) << " - " << A(1); break;
            case Op::Mul: out << A(0) << " * " << A(1); break;
            case Op::Div: out << A(0) << " / " << A(1); break;
            case Op::Neg: out << "-" << A(0); break;
            case Op::Pow: out << "gePow(" << A(0) << ", " << A(1) << ")"; break;
            case Op::Sin: out << "geSin(" << A(0) << ")"; break;
            case Op::Cos: out << "geCos(" << A(0) << ")"; break;
            case Op::Sqrt: out << "geSqrt(" << A(0) << ")"; break;
            case Op::Abs: out << "geAbs(" << A(0) << ")"; break;
            case Op::Exp: out << "geExp(" << A(0) << ")"; break;
            case Op::Log: out << "geLog(" << A(0) << ")"; break;
            case Op::Tan: case Op::Tanh: case Op::Sinh: case Op::Cosh: case Op::Asin: case Op::Acos: case Op::Atan: case Op::Asinh: case Op::Acosh: case Op::Log10:
                out << math1Of(n.op)->device << "(" << A(0) << ")"; break;
            case Op::Less: out << "(" << val(0) << " < " << val(1) << " ? T(1) : T(0))"; break;
            case Op::LessEq: out << "(" << val(0) << " <= " << val(1) << " ? T(1) : T(0))"; break;
            case Op::Greater: out << "(" << val(0) << " > " << val(1) << " ? T(1) : T(0))"; break;
            case Op::GreaterEq: out << "(" << val(0) << " >= " << val(1) << " ? T(1) : T(0))"; break;
            case Op::Eq: out << "(" << val(0) << " == " << val(1) << " ? T(1) : T(0))"; break;
            case Op::Or: out << "((" << val(0) << " != T(0) || " << val(1) << " != T(0)) ? T(1) : T(0))"; break;
            case Op::Not: out << "(" << val(0) << " != T(0) ? T(0) : T(1))"; break;
            case Op::Select: out << "(" << val(0) << " != T(0) ? " << ty << "(" << A(1) << ") : " << ty << "(" << A(2) << "))"; break;
            case Op::Sample: {
                const ImageDecl& im = P.images[n.id];
                auto ptr = [&](int id) { return id < 0 ? std::string("(const ") + ctype(im.type) + "*)0" : std::string("(const ") + ctype(P.images[id].type) + "*)P.img[" + std::to_string(id) + "]"; };
                out << "geSample<T>(" << ptr(n.id) << ", " << ptr(n.off[0]) << ", " << ptr(n.off[1]) << ", (long)P.dim[" << im.dims[0] << "], (long)P.dim[" << im.dims[1] << "], " << im.channels
                    << ", " << n.chan << ", " << A(0) << ", " << A(1) << ")";
                break;
            }
        }
        out << ";\n";
    }
    void prologue(const TermInfo& t, const std::vector<int>& slots) {
        if (t.centered) {
            int nd = (int)t.dims.size();
            for (int k = 0; k < nd; ++k) out << "            const long n" << k << " = (long)P.dim[" << t.dims[k] << "];\n";
            out << "            const long i0 = i % n0;\n";
            if (nd > 1) out << "            const long i1 = (i / n0)" << (nd > 2 ? " % n1" : "") << ";\n";
            if (nd > 2) out << "            const long i2 = i / (n0 * n1);\n";
        } else {
            out << "            const long i0 = i, n0 = 0; (void)i0; (void)n0;\n";
            // a vertex of an nd-dimensional index space is stored as nd interleaved int32 (the reference's Index struct
            // {d0, d1, ...}, o.t:355-387, util.t:677-682); its linear offset is d0 + n0 * (d1 + n1 * d2)
            for (int s : slots) {
                const std::vector<int>& sd = P.slotDims[s];
                const int nd = (int)sd.size();
                out << "            const long e" << s << " = (long)P.gidx[" << s << "][" << nd << " * i]";
                if (nd > 1) out << " + (long)P.dim[" << sd[0] << "] * ((long)P.gidx[" << s << "][" << nd << " * i + 1]";
                if (nd > 2) out << " + (long)P.dim[" << sd[1] << "] * (long)P.gidx[" << s << "][" << nd << " * i + 2]";
                if (nd > 1) out << ")";
                out << ";\n";
            }
        }
    }
};

void collectSlots(const Problem& P, int e, std::set<int>& slots, std::set<int>& seen) {
    if (!seen.insert(e).second) return;
    const Node& n = P.dag.nodes[e];
    if ((n.op == Op::Load || n.op == Op::Unknown) && n.graph) slots.insert(n.slot);
    for (int a : n.args) collectSlots(P, a, slots, seen);
}

CompiledSpec finish(Problem& P) {
    CompiledSpec cs;
    cs.images = P.images; cs.params = P.params; cs.graphs = P.graphs; cs.usePreconditioner = P.usePreconditioner; cs.maxDimIndex = P.maxDim;
    if (P.nUnknown == 0) fail("the specification declares no Unknown");
    if (P.nUnknown > 8) fail("at most 8 unknown arrays are supported");
    if (P.images.size() > 24 || P.params.size() > 40) fail("too many images / parameters for the generic parameter block");
    if (P.terms.empty()) fail("the specification declares no Energy term");
    // ---- classify terms, apply the default bounds rule
    std::vector<TermInfo> infos;
    for (int& term : P.terms) {
        TermInfo t; std::set<int> seen; analyse(P, term, t, seen);
        if (!t.any) { double c; if (P.dag.isConst(term, &c) && c == 0) continue; fail("residual must actually use some image"); }
        if (t.centered && !t.usesBounds) {   // o.t:1930-1934
            Node b; b.op = Op::Bounds; b.nd = (int)t.dims.size();
            for (int k = 0; k < 3; ++k) { b.off[k] = t.bmin[k]; b.off2[k] = t.bmax[k]; }
            bool trivial = true; for (int k = 0; k < 3; ++k) trivial = trivial && t.bmin[k] == 0 && t.bmax[k] == 0;
            if (!trivial) term = P.dag.select(P.dag.add(b), term, P.dag.constant(0));
        }
        if ((int)t.deps.size() > 16) fail("one residual touches more than 16 unknown scalars (" + std::to_string(t.deps.size()) + ")");
        infos.push_back(t);
        GroupDecl g; g.centered = t.centered; g.dims = t.dims; g.graph = t.graph; g.ndeps = std::max<int>(1, (int)t.deps.size());
        g.centerUnknown = 0;
        if (t.centered) for (const ImageDecl& im : P.images) if (im.unknown && im.dims == t.dims) { g.centerUnknown = im.unknownId; break; }
        cs.groups.push_back(g);
    }
    std::vector<int> terms; for (int t : P.terms) { double c; if (!(P.dag.isConst(t, &c) && c == 0)) terms.push_back(t); }
    if (cs.groups.size() > 64) fail("more than 64 scalar residuals per element are not supported");
    // ---- exclude expressions per index space
    std::vector<TermInfo> exInfos;
    for (int e : P.excludes) { TermInfo t; std::set<int> seen; analyse(P, e, t, seen); if (!t.any || !t.centered) fail("Exclude() must be a centred expression over an image"); exInfos.push_back(t); }
    // ---- generate
    Codegen cg(P);
    cg.out << "template <class T>\nstruct GenSpec {\n    typedef GenParams<T> Params;\n";
    cg.out << "    template <class A> __device__ static void eval(const Params& P, int g, long i, A& X, typename A::S* out) {\n        using S = typename A::S;\n        switch (g) {\n";
    for (size_t g = 0; g < terms.size(); ++g) {
        cg.out << "        case " << g << ": {\n";
        std::set<int> slots, s2; collectSlots(P, terms[g], slots, s2);
        cg.prologue(infos[g], std::vector<int>(slots.begin(), slots.end()));
        std::set<int> done; cg.emit(terms[g], done, infos[g], false);
        cg.out << "            out[0] = S(" << cg.v(terms[g]) << ");\n        } break;\n";
    }
    cg.out << "        default: out[0] = S(T(0));\n        }\n    }\n";
    cg.out << "    __device__ static bool excluded(const Params& P, int a, long i) {\n        switch (a) {\n";
    for (const ImageDecl& im : P.images) {
        if (!im.unknown) continue;
        cg.out << "        case " << im.unknownId << ": {\n";
        for (size_t k = 0; k < P.excludes.size(); ++k) {
            if (exInfos[k].dims != im.dims) continue;
            cg.out << "          {\n";
            cg.prologue(exInfos[k], {});
            std::set<int> done; cg.emit(P.excludes[k], done, exInfos[k], true);
            cg.out << "            if (" << cg.v(P.excludes[k]) << " != T(0)) return true;\n          }\n";
        }
        cg.out << "        } break;\n";
    }
    cg.out << "        default: break;\n        }\n        (void)P; (void)i;\n        return false;\n    }\n};\n";
    cs.cudaSource = cg.out.str();
    cs.ok = true;
    return cs;
}

}  // namespace

CompiledSpec compileSpecSource(const std::string& source, const std::string& chunkName) {
    CompiledSpec cs;
    try {
        Parser ps(source);
        std::shared_ptr<Block> chunk = ps.block();
        if (ps.tok.t != Token::Eof) fail("unexpected '" + ps.tok.s + "'", ps.tok.line);
        Problem P;
        Interp I(P);
        installPrelude(I);
        auto top = std::make_shared<Scope>(); top->parent = I.globals;
        try { I.execBlock(*chunk, top); } catch (ReturnEx&) {}
        cs = finish(P);
    } catch (std::exception& e) {
        cs.ok = false; cs.error = chunkName + ": " + e.what();
    }
    return cs;
}

CompiledSpec compileSpecFile(const std::string& filename) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) { CompiledSpec cs; cs.error = "cannot open energy specification '" + filename + "'"; return cs; }
    std::stringstream ss; ss << f.rdbuf();
    return compileSpecSource(ss.str(), filename);
}

}  // namespace fe
}  // namespace optb
