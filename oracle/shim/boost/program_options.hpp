// TEST INFRASTRUCTURE (oracle/): minimal Boost.Program_options stand-in (Boost is
// absent from this image) so /root/reference/src/main.cc compiles unchanged into
// oracle/_ref/.  Supports exactly what main.cc:312-354 uses: string-valued
// options, required(), positional mapping, --name value / --name=value, -h.
#ifndef HMC_ORACLE_SHIM_PROGRAM_OPTIONS_HPP
#define HMC_ORACLE_SHIM_PROGRAM_OPTIONS_HPP
#include <map>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

struct error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct value_semantic {
  bool is_required = false;
  value_semantic *required() {
    is_required = true;
    return this;
  }
};
template <class T> value_semantic *value() { return new value_semantic(); }

struct option_def {
  std::string long_name, short_name, help;
  std::shared_ptr<value_semantic> sem; // null => flag
};

class options_description;
struct options_adder {
  options_description *owner;
  options_adder &operator()(const char *name, const char *help);
  options_adder &operator()(const char *name, value_semantic *s, const char *help);
};

class options_description {
public:
  explicit options_description(const std::string &caption) : caption(caption) {}
  options_adder add_options() { return options_adder{this}; }
  std::string caption;
  std::vector<option_def> opts;
  const option_def *find(const std::string &n) const {
    for (auto &o : opts)
      if (o.long_name == n || (!o.short_name.empty() && o.short_name == n)) return &o;
    return nullptr;
  }
};
inline void split_name(const char *name, option_def &d) {
  std::string s(name);
  auto c = s.find(',');
  d.long_name = s.substr(0, c);
  if (c != std::string::npos) d.short_name = s.substr(c + 1);
}
inline options_adder &options_adder::operator()(const char *name, const char *help) {
  option_def d;
  split_name(name, d);
  d.help = help;
  owner->opts.push_back(d);
  return *this;
}
inline options_adder &options_adder::operator()(const char *name, value_semantic *s,
                                                const char *help) {
  option_def d;
  split_name(name, d);
  d.help = help;
  d.sem.reset(s);
  owner->opts.push_back(d);
  return *this;
}
inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
  os << d.caption << ":\n";
  for (auto &o : d.opts) os << "  --" << o.long_name << "  " << o.help << "\n";
  return os;
}

struct positional_options_description {
  std::vector<std::string> names;
  positional_options_description &add(const char *n, int) {
    names.push_back(n);
    return *this;
  }
};

struct variable_value {
  std::string v;
  template <class T> const T &as() const { return v; }
};

struct variables_map : std::map<std::string, variable_value> {
  const options_description *desc = nullptr;
  size_t count(const std::string &k) const {
    return std::map<std::string, variable_value>::count(k);
  }
  const variable_value &operator[](const std::string &k) const { return at(k); }
};

struct parsed_options {
  std::vector<std::pair<std::string, std::string>> kv;
  const options_description *desc;
};

class command_line_parser {
  std::vector<std::string> args;
  const options_description *desc = nullptr;
  const positional_options_description *pod = nullptr;

public:
  command_line_parser(int argc, char **argv) {
    for (int i = 1; i < argc; i++) args.push_back(argv[i]);
  }
  command_line_parser &options(const options_description &d) {
    desc = &d;
    return *this;
  }
  command_line_parser &positional(const positional_options_description &p) {
    pod = &p;
    return *this;
  }
  parsed_options run() {
    parsed_options out;
    out.desc = desc;
    size_t npos = 0;
    for (size_t i = 0; i < args.size(); i++) {
      const std::string &a = args[i];
      if (a.size() > 1 && a[0] == '-') {
        std::string name = a.substr(a[1] == '-' ? 2 : 1), val;
        bool has_val = false;
        auto eq = name.find('=');
        if (eq != std::string::npos) {
          val = name.substr(eq + 1);
          name = name.substr(0, eq);
          has_val = true;
        }
        const option_def *o = desc->find(name);
        if (!o) throw error("unrecognised option '" + a + "'");
        if (o->sem) {
          if (!has_val) {
            if (i + 1 >= args.size())
              throw error("the required argument for option '--" + o->long_name +
                          "' is missing");
            val = args[++i];
          }
        }
        out.kv.emplace_back(o->long_name, val);
      } else {
        if (!pod || npos >= pod->names.size())
          throw error("too many positional options have been specified");
        out.kv.emplace_back(pod->names[npos++], a);
      }
    }
    return out;
  }
};

inline void store(const parsed_options &p, variables_map &vm) {
  vm.desc = p.desc;
  for (auto &kv : p.kv)
    if (!vm.count(kv.first)) vm.emplace(kv.first, variable_value{kv.second});
}
inline void notify(variables_map &vm) {
  if (!vm.desc) return;
  for (auto &o : vm.desc->opts)
    if (o.sem && o.sem->is_required && !vm.count(o.long_name))
      throw error("the option '--" + o.long_name + "' is required but missing");
}

} // namespace program_options
} // namespace boost
#endif
