// psy_codegen.cpp — equation strings -> CUDA translation unit, behind the C ABI (psy_emit_model_source & co).
//
// GPU counterpart of the reference's model assembly: newPsydiff pastes prepareEquations(latentEquations) /
// prepareEquations(manifestEquations) into modelTemplate() by placeholder replacement (R/prepareModel.R:479-512,
// 737-770; template R/modelTemplate.R:5-1011).  There the template is Rcpp/Armadillo/odeint C++ and every identifier that
// names an entry of `parameters` is bound to an arma::mat / double pulled out of an Rcpp::List by string key on every
// call.  Here the template is a model struct for psy_filter.cuh; a referenced container becomes a fixed-size
// psy::Mat<r,c> (or double) whose free entries read the instance's slot array psy_p_[k] and whose fixed entries are
// literal constants, so structural zeros fold away at compile time and nothing is looked up by name.
//
// Living in the shared library (not in a host-language package) is what lets ANY host — the R .Call shim of
// r/psydiff_b200_shim.c, the Python mirror in psydiff_b200/codegen.py — produce the same source from the same model.
#include "../../include/psydiff_b200.h"

#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <regex>
#include <set>
#include <string>
#include <vector>

int psy_set_error(int code, const char* msg);  // psy_capi.cu: fills psy_last_error()

namespace {

struct Cont {
  std::string name;
  bool scalar;
  int rows, cols;
  const double* values;  // column-major
  const int* slots;      // column-major, -1 = fixed
  double val(int i, int j) const { return values[i + (size_t)j * rows]; }
  int slot(int i, int j) const { return slots[i + (size_t)j * rows]; }
};

int err(const std::string& m) { return psy_set_error(PSY_ERR_ARG, m.c_str()); }

bool is_ident_char(char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_'; }

const Cont* find(const std::vector<Cont>& cs, const std::string& name) {
  for (const auto& c : cs)
    if (c.name == name) return &c;
  return nullptr;
}

// identifiers of the equation text that name a parameter container, in order of first appearance
// (prepareEquations, R/prepareModel.R:737-766)
std::vector<const Cont*> referenced(const std::string& eq, const std::vector<Cont>& cs) {
  std::vector<const Cont*> out;
  size_t i = 0;
  while (i < eq.size()) {
    if (!is_ident_char(eq[i])) { i++; continue; }
    size_t j = i;
    while (j < eq.size() && is_ident_char(eq[j])) j++;
    const Cont* c = find(cs, eq.substr(i, j - i));
    if (c) {
      bool seen = false;
      for (auto* o : out) seen = seen || (o == c);
      if (!seen) out.push_back(c);
    }
    i = j;
  }
  return out;
}

// A double as a C++ literal that reads back bit for bit: shortest round-trip digits; fixed notation for decimal
// exponents in [-4, 16), otherwise d.ddde±XX
std::string lit(double v) {
  if (std::isnan(v)) return "(0.0/0.0)";
  if (std::isinf(v)) return v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
  std::string sign = std::signbit(v) ? "-" : "";
  v = std::fabs(v);
  if (v == 0.0) return sign + "0.0";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
  std::string s(buf, r.ptr);                       // d[.ddd]e±XX
  const size_t e = s.find('e');
  std::string digits;
  for (size_t k = 0; k < e; k++)
    if (s[k] != '.') digits += s[k];
  const int decpt = std::stoi(s.substr(e + 1)) + 1;  // value = 0.DIGITS x 10^decpt
  const int nd = (int)digits.size();
  if (decpt <= -4 || decpt > 16) {
    std::string out = sign + digits.substr(0, 1);
    if (nd > 1) out += "." + digits.substr(1);
    const int ex = decpt - 1;
    char eb[16];
    snprintf(eb, sizeof eb, "e%c%02d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
    return out + eb;
  }
  if (decpt <= 0) return sign + "0." + std::string((size_t)(-decpt), '0') + digits;
  if (decpt >= nd) return sign + digits + std::string((size_t)(decpt - nd), '0') + ".0";
  return sign + digits.substr(0, (size_t)decpt) + "." + digits.substr((size_t)decpt);
}

std::string elem(const Cont& c, int i, int j) {
  const int k = c.slot(i, j);
  return k >= 0 ? "psy_p_[" + std::to_string(k) + "]" : lit(c.val(i, j));
}

// declaration of one container inside an equation function
std::string bind(const Cont& c) {
  if (c.scalar) return "    const double " + c.name + " = " + elem(c, 0, 0) + ";\n";
  std::string out = "    psy::Mat<" + std::to_string(c.rows) + "," + std::to_string(c.cols) + "> " + c.name + ";\n";
  for (int j = 0; j < c.cols; j++)
    for (int i = 0; i < c.rows; i++)
      out += "    " + c.name + ".a[" + std::to_string(i + j * c.rows) + "] = " + elem(c, i, j) + ";\n";
  return out;
}

// (diagonal, lower) by STRUCTURE: an entry counts as zero only if it is fixed at exactly 0
void structure(const Cont& c, bool& diag, bool& lower) {
  diag = lower = true;
  for (int i = 0; i < c.rows; i++)
    for (int j = 0; j < c.cols; j++) {
      if (i == j) continue;
      if (c.slot(i, j) >= 0 || c.val(i, j) != 0.0) {
        diag = false;
        if (j > i) lower = false;
      }
    }
}

std::string indent_statements(const std::string& eq) {
  size_t a = 0, b = eq.size();
  while (a < b && isspace((unsigned char)eq[a])) a++;
  while (b > a && isspace((unsigned char)eq[b - 1])) b--;
  std::string out = "    ";
  for (size_t k = a; k < b; k++) {
    out += eq[k];
    if (eq[k] == '\n') out += "    ";
  }
  return out + "\n";
}

// Initialised Armadillo temporaries (`arma::mat T = arma::expm(A);`, vignettes/psydiff-basics.Rmd:94-98) get the type of their
// initialiser: on the device every matrix has compile-time extents, so the declared type becomes `auto`.
std::string device_statements(const std::string& eq) {
  static const std::regex typed(R"(\barma\s*::\s*(?:mat|colvec|vec|rowvec)\s+([A-Za-z_][A-Za-z0-9_]*)\s*=(?!=))");
  return std::regex_replace(eq, typed, "auto $1 =");
}

// prepareEquations (R/prepareModel.R:737-770): declarations of the containers the text names, then the statements
std::string prepare_equations(const std::string& eq, const std::vector<Cont>& cs) {
  std::string out;
  for (auto* c : referenced(eq, cs)) out += bind(*c);
  return out + indent_statements(device_statements(eq));
}

std::string strip(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

const std::set<std::string>& math_ok() {
  static const std::set<std::string> s = {"exp", "log", "sqrt", "pow", "sin", "cos", "tan", "tanh", "sinh", "cosh", "fabs", "abs",
                                          "log1p", "expm1", "atan", "std", "t", "person"};
  return s;
}

// Sparsity of the drift's Jacobian as far as it can be read off the equation text: bit r of mask[i] is set if dx_i MAY
// depend on x_r.  Conservative: anything that is not a plain `dx(i) = <expression in x(k), scalar parameters, t, person>;`
// statement (or a single `dx = MATRIX*x;` with structurally zero entries) makes every component depend on every state.
// The kernel skips the sigma-point columns that cannot change a component; there f(x+) = f(x-) = f(x0) bit for bit.
std::vector<unsigned> drift_dependency(int n, const std::string& latent, const std::vector<Cont>& cs) {
  const unsigned full = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
  const std::vector<unsigned> all((size_t)n, full);
  std::string flat = latent;
  for (auto& ch : flat)
    if (ch == '\n') ch = ' ';
  std::vector<std::string> stmts;
  {
    size_t a = 0;
    while (a <= flat.size()) {
      size_t b = flat.find(';', a);
      if (b == std::string::npos) b = flat.size();
      std::string st = strip(flat.substr(a, b - a));
      if (!st.empty()) stmts.push_back(st);
      a = b + 1;
    }
  }
  static const std::regex elem_stmt(R"(^dx\s*[\(\[]\s*(\d+)\s*[\)\]]\s*=(?!=)\s*([\s\S]+)$)");
  static const std::regex x_ref(R"(\bx\s*[\(\[]\s*(\d+)\s*[\)\]])");
  static const std::regex ident(R"([A-Za-z_][A-Za-z0-9_]*)");
  static const std::regex whole(R"(^dx\s*=(?!=)\s*([A-Za-z_][A-Za-z0-9_]*)\s*\*\s*x$)");
  std::vector<long long> mask((size_t)n, -1);
  for (const auto& st : stmts) {
    std::smatch mm;
    if (std::regex_match(st, mm, elem_stmt)) {
      const std::string idx = mm[1].str();
      if (idx.size() > 6) return all;
      const int i = std::stoi(idx);
      if (i >= n || mask[(size_t)i] >= 0) return all;  // out of range / assigned twice: give up
      const std::string expr = mm[2].str();
      unsigned deps = 0;
      std::string rest;
      {
        auto it = std::sregex_iterator(expr.begin(), expr.end(), x_ref), end = std::sregex_iterator();
        size_t pos = 0;
        for (; it != end; ++it) {
          const std::string k = (*it)[1].str();
          if (k.size() > 6 || std::stoi(k) >= n) return all;
          deps |= 1u << std::stoi(k);
          rest += expr.substr(pos, (size_t)it->position() - pos) + " ";
          pos = (size_t)it->position() + (size_t)it->length();
        }
        rest += expr.substr(pos);
      }
      for (auto it = std::sregex_iterator(rest.begin(), rest.end(), ident); it != std::sregex_iterator(); ++it) {
        const std::string w = it->str();
        if (w == "e" || w == "E") continue;  // exponent of a float literal
        if (math_ok().count(w)) continue;
        const Cont* c = find(cs, w);
        if (c && c->scalar) continue;
        return all;  // x used whole, dx on the right, matrices, temporaries, ...
      }
      mask[(size_t)i] = deps;
      continue;
    }
    if (std::regex_match(st, mm, whole) && stmts.size() == 1) {
      const Cont* c = find(cs, mm[1].str());
      if (c && !c->scalar && c->rows == n && c->cols == n) {
        std::vector<unsigned> out((size_t)n, 0u);
        for (int i = 0; i < n; i++)
          for (int k = 0; k < n; k++)
            if (c->slot(i, k) >= 0 || c->val(i, k) != 0.0) out[(size_t)i] |= 1u << k;
        return out;
      }
    }
    return all;
  }
  std::vector<unsigned> out((size_t)n);
  for (int i = 0; i < n; i++) out[(size_t)i] = mask[(size_t)i] < 0 ? 0u : (unsigned)mask[(size_t)i];  // never assigned: dx stays 0
  return out;
}

std::string model_template(bool sr) {
  return std::string("// generated by psydiff_b200.codegen — model translation unit for sm_100a\n"
                     "#include \"psy_filter.cuh\"\n\n"
                     "struct PsyUserModel {\n"
                     "  static constexpr int N = NLATENTPLACEHOLDER, M = NMANIFESTPLACEHOLDER, NSLOT = NSLOTPLACEHOLDER, NFREE = NFREEPLACEHOLDER;\n"
                     "  static constexpr int L_STRUCT = LSTRUCTPLACEHOLDER;\n"
                     "  static constexpr bool R_DIAG = RDIAGPLACEHOLDER;\n"
                     "  static constexpr bool SR = ") +
         (sr ? "true" : "false") +
         ";\n"
         "  static constexpr int STEPPER = STEPPERPLACEHOLDER;\n"
         "  // bit r of dep(i): drift component i may read x_r (Jacobian sparsity read off the equation text)\n"
         "  static __host__ __device__ constexpr unsigned dep(int i) { return DEPPLACEHOLDER; }\n\n"
         "LATENTEQUATIONPLACEHOLDER\n"
         "MEASUREMENTEQUATIONPLACEHOLDER\n"
         "INITIALPLACEHOLDER"
         "};\n\nPSY_INSTANTIATE(PsyUserModel)\n";
}

void replace_all(std::string& s, const std::string& key, const std::string& val) {  // stringr::str_replace_all, R/prepareModel.R:511-512
  size_t pos = 0;
  while ((pos = s.find(key, pos)) != std::string::npos) {
    s.replace(pos, key.size(), val);
    pos += val.size();
  }
}

int copy_out(const std::string& s, char* out, size_t cap, size_t* needed) {
  if (needed) *needed = s.size() + 1;
  if (!out || cap < s.size() + 1) {
    if (needed && !out) return PSY_OK;  // size query
    return err("output buffer too small for the generated source (" + std::to_string(s.size() + 1) + " bytes needed)");
  }
  memcpy(out, s.c_str(), s.size() + 1);
  return PSY_OK;
}

int gather(const psy_container* cs, int nc, std::vector<Cont>& out) {
  if (nc < 0 || (nc > 0 && !cs)) return err("containers: bad arguments");
  for (int k = 0; k < nc; k++) {
    const psy_container& c = cs[k];
    if (!c.name || !c.values || !c.slots || c.rows < 1 || c.cols < 1) return err("container " + std::to_string(k) + ": bad arguments");
    if (c.is_scalar && (c.rows != 1 || c.cols != 1)) return err(std::string("container ") + c.name + ": a scalar is 1 x 1");
    out.push_back(Cont{c.name, c.is_scalar != 0, c.rows, c.cols, c.values, c.slots});
  }
  return PSY_OK;
}

std::string hexu(unsigned v) {
  char b[16];
  snprintf(b, sizeof b, "0x%xu", v);
  return b;
}

}  // namespace

extern "C" {

int psy_drift_dependency(int n, const char* latent_equations, const psy_container* containers, int n_containers, unsigned* mask) {
  if (n < 1 || n > 31 || !latent_equations || !mask) return err("psy_drift_dependency: bad arguments");
  std::vector<Cont> cs;
  int rc = gather(containers, n_containers, cs);
  if (rc) return rc;
  const auto d = drift_dependency(n, latent_equations, cs);
  for (int i = 0; i < n; i++) mask[i] = d[(size_t)i];
  return PSY_OK;
}

int psy_model_template(int sr_update, char* out, size_t cap, size_t* needed) {
  return copy_out(model_template(sr_update != 0), out, cap, needed);
}

int psy_emit_model_source(int n, int m, int n_slots, const char* latent_equations, const char* manifest_equations,
                          const psy_container* containers, int n_containers, int sr_update, int stepper, char* out, size_t cap,
                          size_t* needed) {
  if (n < 1 || n > 31 || m < 1 || n_slots < 0 || !latent_equations || !manifest_equations)
    return err("psy_emit_model_source: bad arguments");
  if (stepper != PSY_STEPPER_RK4 && stepper != PSY_STEPPER_DOPRI5) return err("psy_emit_model_source: stepper must be PSY_STEPPER_RK4 or PSY_STEPPER_DOPRI5");
  std::vector<Cont> cs;
  int rc = gather(containers, n_containers, cs);
  if (rc) return rc;
  const Cont *A0 = find(cs, "A0LogChol"), *Ll = find(cs, "LLogChol"), *Rl = find(cs, "RLogChol"), *m0 = find(cs, "m0");
  if (!A0 || !Ll || !Rl || !m0) return err("containers must include A0LogChol, LLogChol, RLogChol and m0 (R/prepareModel.R:445-459)");
  if (A0->rows != n || A0->cols != n || Ll->rows != n || Ll->cols != n) return err("A0 and L must be nlatent x nlatent");
  if (Rl->rows != m || Rl->cols != m) return err("Rchol must be nmanifest x nmanifest");
  if (m0->rows * m0->cols != n) return err("m0 must have nlatent elements");
  for (const auto& c : cs)
    for (int j = 0; j < c.cols; j++)
      for (int i = 0; i < c.rows; i++)
        if (c.slot(i, j) >= n_slots) return err("container " + c.name + ": slot id out of range");
  bool dg, lower, l_diag, r_diag;
  structure(*A0, dg, lower);
  if (!lower)
    return err("A0 must be lower triangular (it is used as the Cholesky root of the initial covariance); "
               "entries above the diagonal are unsupported on device");
  structure(*Ll, l_diag, lower);
  structure(*Rl, r_diag, lower);
  const auto dep = drift_dependency(n, latent_equations, cs);
  std::string chain;
  for (int i = 0; i + 1 < n; i++) chain += (i ? " : " : "") + ("i == " + std::to_string(i) + " ? " + hexu(dep[(size_t)i]));
  chain += (n > 1 ? " : " : "") + hexu(dep[(size_t)n - 1]);
  // latentEquation (R/prepareModel.R:493-505) / measurementEquation (R/prepareModel.R:479-491)
  const std::string latent = "  static __device__ __forceinline__ void latentEquation(const psy::Vec<N>& x, psy::Vec<N>& dx,\n"
                             "      const double (&psy_p_)[NFREE], const int person, const double t) {\n" +
                             prepare_equations(latent_equations, cs) + "  }\n";
  const std::string manifest = "  static __device__ __forceinline__ void measurementEquation(const psy::Vec<N>& x, psy::Vec<M>& y,\n"
                               "      const double (&psy_p_)[NFREE], const int person, const double t) {\n" +
                               prepare_equations(manifest_equations, cs) + "  }\n";
  // initial state: m0, A0LogChol, LLogChol as stored (log-Cholesky; the kernel applies logChol2Chol_C); Rchol = logChol2Chol(RLogChol)
  std::string ini = "  static __device__ __forceinline__ void initial(double (&m0)[N], double (&A0lc)[N * N], double (&Llc)[N * N],\n"
                    "      const double (&psy_p_)[NFREE]) {\n";
  for (int i = 0; i < n; i++) {  // m0 in column-major element order
    const int k = m0->slots[i];
    ini += "    m0[" + std::to_string(i) + "] = " + (k >= 0 ? "psy_p_[" + std::to_string(k) + "]" : lit(m0->values[i])) + ";\n";
  }
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) ini += "    A0lc[" + std::to_string(i + j * n) + "] = " + elem(*A0, i, j) + ";\n";
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) ini += "    Llc[" + std::to_string(i + j * n) + "] = " + elem(*Ll, i, j) + ";\n";
  ini += "  }\n\n";
  ini += "  static __device__ __forceinline__ void rchol(double (&Rc)[M * M], const double (&psy_p_)[NFREE]) {\n";
  for (int j = 0; j < m; j++)
    for (int i = 0; i < m; i++) {
      const std::string e = elem(*Rl, i, j);
      ini += "    Rc[" + std::to_string(i + j * m) + "] = " + (i == j ? "exp(" + e + ")" : e) + ";\n";
    }
  ini += "  }\n";
  std::string code = model_template(sr_update != 0);
  replace_all(code, "NLATENTPLACEHOLDER", std::to_string(n));
  replace_all(code, "NMANIFESTPLACEHOLDER", std::to_string(m));
  replace_all(code, "NSLOTPLACEHOLDER", std::to_string(n_slots));
  replace_all(code, "NFREEPLACEHOLDER", std::to_string(n_slots > 1 ? n_slots : 1));
  replace_all(code, "LSTRUCTPLACEHOLDER", l_diag ? "psy::L_DIAG" : "psy::L_FULL");
  replace_all(code, "RDIAGPLACEHOLDER", r_diag ? "true" : "false");
  replace_all(code, "STEPPERPLACEHOLDER", std::to_string(stepper));
  replace_all(code, "DEPPLACEHOLDER", chain);
  replace_all(code, "MEASUREMENTEQUATIONPLACEHOLDER", manifest);
  replace_all(code, "LATENTEQUATIONPLACEHOLDER", latent);
  replace_all(code, "INITIALPLACEHOLDER", ini);
  return copy_out(code, out, cap, needed);
}

}  // extern "C"
