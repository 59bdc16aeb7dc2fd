// crixus_b200 -- .ini front end.  A C++ reader with the rules (and, in ini.cpp, the parse-loop structure -- BSD notice there) of the reference's vendored inih
// (/root/reference/src/ini/ini.c:60-167, src/ini/cpp/INIReader.cpp): sections, name=value or name:value, ';'/'#'
// comment lines, ';' inline comments only after whitespace, continuation lines (leading whitespace), lines cut at
// 199 characters, case-insensitive section.name keys, repeated keys joined with '\n', strtol/strtod value parsing,
// booleans true/yes/on/1 and false/no/off/0.
#pragma once
#include <map>
#include <string>

class CxIni {
 public:
  explicit CxIni(const std::string &filename);
  int ParseError() const { return error_; }  // -1: cannot open; >0: first bad line; 0: ok
  std::string Get(const std::string &section, const std::string &name, const std::string &def) const;
  long GetInteger(const std::string &section, const std::string &name, long def) const;
  double GetReal(const std::string &section, const std::string &name, double def) const;
  bool GetBoolean(const std::string &section, const std::string &name, bool def) const;

 private:
  static std::string MakeKey(const std::string &section, const std::string &name);
  int error_;
  std::map<std::string, std::string> values_;
};
