// host/tpt_ini.h -- config.ini reader with the call surface main() uses from inipp
// (third_party/inipp.h:73-91 `extract`, :93-159 `Ini::sections / errors / parse / generate`).
// Own implementation of that subset (no ${} interpolation, which the driver never uses):
//   * lines are trimmed; empty lines and lines starting with ';' are skipped
//   * "[name]" opens a section; "key = value" splits at the first '='; the first definition of
//     a key wins, a repeated key is recorded in `errors`
//   * extract(text, dst) parses the WHOLE string as T (bools as true/false) and leaves dst
//     untouched on failure -- so a missing key keeps the caller's default (main.cpp:29-56)
//   * generate(os) echoes "[section]" / "key=value" lines in map order (main.cpp:58)
#ifndef TPT_HOST_INI_H_
#define TPT_HOST_INI_H_

#include <cctype>
#include <istream>
#include <list>
#include <map>
#include <ostream>
#include <sstream>
#include <string>

namespace inipp {

template <typename CharT, typename T>
bool extract(const std::basic_string<CharT> &text, T &dst) {
  std::basic_istringstream<CharT> in(text);
  T parsed;
  CharT trailing;
  if (!(in >> std::boolalpha >> parsed)) return false;
  if (in >> trailing) return false; // something left after the value
  dst = parsed;
  return true;
}
template <typename CharT>
bool extract(const std::basic_string<CharT> &text, std::basic_string<CharT> &dst) {
  dst = text;
  return true;
}

template <class CharT> class Ini {
public:
  typedef std::basic_string<CharT> String;
  typedef std::map<String, String> Section;
  typedef std::map<String, Section> Sections;

  Sections sections;
  std::list<String> errors;

  void parse(std::basic_istream<CharT> &in) {
    String raw, current;
    while (std::getline(in, raw)) {
      String line = trimmed(raw);
      if (line.empty() || line[0] == CharT(';')) continue;
      if (line[0] == CharT('[')) {
        if (line[line.size() - 1] == CharT(']'))
          current = line.substr(1, line.size() - 2);
        else
          errors.push_back(line);
        continue;
      }
      typename String::size_type eq = line.find(CharT('='));
      if (eq == String::npos || eq == 0) {
        errors.push_back(line);
        continue;
      }
      String key = trimmed(line.substr(0, eq));
      String value = trimmed(line.substr(eq + 1));
      Section &sec = sections[current];
      if (!sec.insert(std::make_pair(key, value)).second) errors.push_back(line);
    }
  }

  void generate(std::basic_ostream<CharT> &os) const {
    for (typename Sections::const_iterator s = sections.begin(); s != sections.end(); ++s) {
      os << CharT('[') << s->first << CharT(']') << std::endl;
      for (typename Section::const_iterator kv = s->second.begin(); kv != s->second.end(); ++kv)
        os << kv->first << CharT('=') << kv->second << std::endl;
    }
  }

  void clear() {
    sections.clear();
    errors.clear();
  }

private:
  static String trimmed(const String &s) {
    typename String::size_type b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) ++b;
    while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
  }
};

} // namespace inipp
#endif
