// crixus_b200 -- .ini reader.  The parse loop below follows inih (the parser vendored by the reference under src/ini/, which
// Crixus's option semantics are defined by): same line rules, same helper structure (rstrip / lskip / find_char_or_comment),
// same limits.  inih is third-party code under the New BSD license; its notice is reproduced here as that license requires:
//
//   inih -- simple .INI file parser.  Copyright (c) 2009, Brush Technology.  All rights reserved.
//   Redistribution and use in source and binary forms, with or without modification, are permitted provided that the
//   following conditions are met: * Redistributions of source code must retain the above copyright notice, this list of
//   conditions and the following disclaimer. * Redistributions in binary form must reproduce the above copyright notice, this
//   list of conditions and the following disclaimer in the documentation and/or other materials provided with the
//   distribution. * Neither the name of Brush Technology nor the names of its contributors may be used to endorse or promote
//   products derived from this software without specific prior written permission.
//   THIS SOFTWARE IS PROVIDED BY BRUSH TECHNOLOGY ''AS IS'' AND ANY EXPRESS OR IMPLIED WARRANTIES, INCLUDING, BUT NOT LIMITED
//   TO, THE IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A PARTICULAR PURPOSE ARE DISCLAIMED. IN NO EVENT SHALL BRUSH
//   TECHNOLOGY BE LIABLE FOR ANY DIRECT, INDIRECT, INCIDENTAL, SPECIAL, EXEMPLARY, OR CONSEQUENTIAL DAMAGES (INCLUDING, BUT NOT
//   LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS OR SERVICES; LOSS OF USE, DATA, OR PROFITS; OR BUSINESS INTERRUPTION) HOWEVER
//   CAUSED AND ON ANY THEORY OF LIABILITY, WHETHER IN CONTRACT, STRICT LIABILITY, OR TORT (INCLUDING NEGLIGENCE OR OTHERWISE)
//   ARISING IN ANY WAY OUT OF THE USE OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF SUCH DAMAGE.
#include "ini.h"
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
const int kMaxLine = 200, kMaxSection = 50, kMaxName = 50;  // ini.h:70, ini.c:17-18
char *rstrip(char *s)
{
  char *p = s + strlen(s);
  while (p > s && isspace((unsigned char)(*--p))) *p = '\0';
  return s;
}
char *lskip(char *s)
{
  while (*s && isspace((unsigned char)(*s))) s++;
  return s;
}
// first occurrence of c, or of a ';' that follows whitespace (ini.c:43-53)
char *find_char_or_comment(char *s, char c)
{
  int was_ws = 0;
  while (*s && *s != c && !(was_ws && *s == ';')) {
    was_ws = isspace((unsigned char)(*s));
    s++;
  }
  return s;
}
}  // namespace

std::string CxIni::MakeKey(const std::string &section, const std::string &name)
{
  std::string key = section + "." + name;
  std::transform(key.begin(), key.end(), key.begin(), ::tolower);
  return key;
}

CxIni::CxIni(const std::string &filename) : error_(0)
{
  FILE *f = fopen(filename.c_str(), "r");
  if (!f) { error_ = -1; return; }
  char line[kMaxLine];
  std::string section, prev_name;
  int lineno = 0;
  auto store = [&](const std::string &name, const char *value) {
    std::string key = MakeKey(section, name);
    std::string &v = values_[key];
    if (!v.empty()) v += "\n";
    v += value;
  };
  while (fgets(line, kMaxLine, f)) {
    lineno++;
    char *start = line;
    if (lineno == 1 && (unsigned char)start[0] == 0xEF && (unsigned char)start[1] == 0xBB && (unsigned char)start[2] == 0xBF) start += 3;
    start = lskip(rstrip(start));
    if (*start == ';' || *start == '#') {
      // comment line
    } else if (!prev_name.empty() && *start && start > line) {
      store(prev_name, start);  // continuation of the previous value
    } else if (*start == '[') {
      char *end = find_char_or_comment(start + 1, ']');
      if (*end == ']') {
        *end = '\0';
        section = std::string(start + 1).substr(0, kMaxSection - 1);
        prev_name.clear();
      } else if (!error_) error_ = lineno;
    } else if (*start && *start != ';') {
      char *end = find_char_or_comment(start, '=');
      if (*end != '=') end = find_char_or_comment(start, ':');
      if (*end == '=' || *end == ':') {
        *end = '\0';
        char *name = rstrip(start);
        char *value = lskip(end + 1);
        end = find_char_or_comment(value, '\0');
        if (*end == ';') *end = '\0';
        rstrip(value);
        prev_name = std::string(name).substr(0, kMaxName - 1);
        store(name, value);
      } else if (!error_) error_ = lineno;
    }
  }
  fclose(f);
}

std::string CxIni::Get(const std::string &section, const std::string &name, const std::string &def) const
{
  auto it = values_.find(MakeKey(section, name));
  return it != values_.end() ? it->second : def;
}
long CxIni::GetInteger(const std::string &section, const std::string &name, long def) const
{
  std::string v = Get(section, name, "");
  char *end;
  long n = strtol(v.c_str(), &end, 0);
  return end > v.c_str() ? n : def;
}
double CxIni::GetReal(const std::string &section, const std::string &name, double def) const
{
  std::string v = Get(section, name, "");
  char *end;
  double n = strtod(v.c_str(), &end);
  return end > v.c_str() ? n : def;
}
bool CxIni::GetBoolean(const std::string &section, const std::string &name, bool def) const
{
  std::string v = Get(section, name, "");
  std::transform(v.begin(), v.end(), v.begin(), ::tolower);
  if (v == "true" || v == "yes" || v == "on" || v == "1") return true;
  if (v == "false" || v == "no" || v == "off" || v == "0") return false;
  return def;
}
