// bci-executor-server — native stand-in for the reference's in-pod executor server
// (executor/server.rs), speaking the same wire contract so an unmodified
// KubernetesCodeExecutor (kubernetes_code_executor.py:95-149) can drive it:
//
//   PUT  /workspace/{path}  body -> file, parents created            -> 204   (server.rs:69-88)
//   GET  /workspace/{path}  file -> application/octet-stream          -> 200   (server.rs:90-96)
//   POST /execute  {"source_code": str, "timeout"?: u64 s, "env"?: {str: str}}
//        -> {"stdout", "stderr", "exit_code", "files": [absolute paths]}        (server.rs:120-179)
//
// Semantics kept: default timeout 60 s; timeout -> ("", "Execution timed out", -1); death by
// signal -> exit_code -1; stdout/stderr decoded lossily (U+FFFD); `files` = regular files directly
// in the workspace whose ctime is later than the start of the request (non-recursive).
// Environment: APP_LISTEN_ADDR (default 0.0.0.0:8000), APP_WORKSPACE (default /workspace).
// What differs, on purpose: the script runs in a fork of the warm zygote (BCI_ZYGOTE_SOCKET,
// bee_code_interpreter_b200/zygote.py — numpy view + broker client preloaded, array work goes to
// the node's GPU broker) or, without one, as `$BCI_PYTHON script.py` with the import hook on
// PYTHONPATH; there is no `upm guess` / `pip install` step (no network in the sandbox image), and
// `..` path components are refused (400) instead of escaping the workspace.
//
// Plain C++17 + POSIX; no CUDA in this process.  One thread per connection, HTTP/1.1 keep-alive,
// Content-Length and chunked request bodies, `Expect: 100-continue`.
#include <arpa/inet.h>
#include <dirent.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <signal.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/types.h>
#include <sys/un.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Config {
  std::string host = "0.0.0.0";
  int port = 8000;
  std::string workspace = "/workspace";
  std::string zygote_socket;   // empty -> exec mode
  std::string python = "python3";
  std::string broker_socket;   // forwarded to the script (BCI_BROKER_SOCKET)
  std::string scratch = "/tmp";
  std::string report_as;       // prefix under which changed files are reported (default: the workspace path)
};
Config g_cfg;

// ------------------------------------------------------------------------------------- JSON
// Just enough JSON for the two message shapes above.
struct JsonValue {
  enum Kind { Null, Bool, Number, String, Object, Array } kind = Null;
  bool b = false;
  double num = 0;
  bool num_is_uint = false;
  unsigned long long u = 0;
  std::string str;
  std::vector<std::pair<std::string, JsonValue>> obj;
  std::vector<JsonValue> arr;
  const JsonValue* get(const char* key) const {
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct JsonParser {
  const char* p;
  const char* end;
  std::string err;
  int depth = 0;

  void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
  bool fail(const char* m) { if (err.empty()) err = m; return false; }

  static void put_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
  }
  bool hex4(uint32_t& v) {
    if (end - p < 4) return fail("truncated \\u escape");
    v = 0;
    for (int i = 0; i < 4; ++i) {
      char c = *p++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else return fail("bad \\u escape");
    }
    return true;
  }
  bool string(std::string& out) {
    if (p >= end || *p != '"') return fail("expected string");
    ++p;
    while (p < end) {
      unsigned char c = (unsigned char)*p++;
      if (c == '"') return true;
      if (c < 0x20) return fail("control character in string");
      if (c != '\\') { out += (char)c; continue; }
      if (p >= end) break;
      char e = *p++;
      switch (e) {
        case '"': out += '"'; break;
        case '\\': out += '\\'; break;
        case '/': out += '/'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'n': out += '\n'; break;
        case 'r': out += '\r'; break;
        case 't': out += '\t'; break;
        case 'u': {
          uint32_t cp;
          if (!hex4(cp)) return false;
          if (cp >= 0xD800 && cp <= 0xDBFF && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
            const char* save = p;
            p += 2;
            uint32_t lo;
            if (!hex4(lo)) return false;
            if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else { p = save; cp = 0xFFFD; }
          } else if (cp >= 0xD800 && cp <= 0xDFFF) {
            cp = 0xFFFD;  // lone surrogate
          }
          put_utf8(out, cp);
          break;
        }
        default: return fail("bad escape");
      }
    }
    return fail("unterminated string");
  }
  bool value(JsonValue& v) {
    if (++depth > 64) return fail("nesting too deep");
    ws();
    if (p >= end) return fail("unexpected end");
    bool ok = true;
    char c = *p;
    if (c == '"') { v.kind = JsonValue::String; ok = string(v.str); }
    else if (c == '{') {
      v.kind = JsonValue::Object;
      ++p; ws();
      if (p < end && *p == '}') { ++p; }
      else for (;;) {
        ws();
        std::string k;
        if (!string(k)) { ok = false; break; }
        ws();
        if (p >= end || *p != ':') { ok = fail("expected ':'"); break; }
        ++p;
        JsonValue child;
        if (!value(child)) { ok = false; break; }
        v.obj.emplace_back(std::move(k), std::move(child));
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == '}') { ++p; break; }
        ok = fail("expected ',' or '}'");
        break;
      }
    } else if (c == '[') {
      v.kind = JsonValue::Array;
      ++p; ws();
      if (p < end && *p == ']') { ++p; }
      else for (;;) {
        JsonValue child;
        if (!value(child)) { ok = false; break; }
        v.arr.push_back(std::move(child));
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == ']') { ++p; break; }
        ok = fail("expected ',' or ']'");
        break;
      }
    } else if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JsonValue::Bool; v.b = true; p += 4; }
    else if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JsonValue::Bool; v.b = false; p += 5; }
    else if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = JsonValue::Null; p += 4; }
    else if (c == '-' || (c >= '0' && c <= '9')) {
      const char* s = p;
      bool integral = c != '-';
      if (*p == '-') ++p;
      while (p < end && *p >= '0' && *p <= '9') ++p;
      if (p < end && (*p == '.' || *p == 'e' || *p == 'E')) {
        integral = false;
        while (p < end && (*p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-' || (*p >= '0' && *p <= '9'))) ++p;
      }
      std::string t(s, p);
      v.kind = JsonValue::Number;
      v.num = strtod(t.c_str(), nullptr);
      if (integral && t.size() <= 19) { v.num_is_uint = true; v.u = strtoull(t.c_str(), nullptr, 10); }
    } else ok = fail("unexpected character");
    --depth;
    return ok;
  }
};

// Appends `s` as a JSON string.  Valid UTF-8 passes through; every byte that is not part of a
// valid sequence becomes U+FFFD — String::from_utf8_lossy of server.rs:163-164.
void json_string(std::string& out, const std::string& s) {
  static const char* hexd = "0123456789abcdef";
  out += '"';
  const unsigned char* p = (const unsigned char*)s.data();
  const unsigned char* end = p + s.size();
  while (p < end) {
    unsigned char c = *p;
    if (c < 0x80) {
      switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        default:
          if (c < 0x20) { out += "\\u00"; out += hexd[c >> 4]; out += hexd[c & 15]; }
          else out += (char)c;
      }
      ++p;
      continue;
    }
    // Well-formed UTF-8 (Unicode table 3-7); an ill-formed sequence is replaced as ONE U+FFFD per
    // maximal valid prefix, which is what from_utf8_lossy (and Python's "replace") do.
    int need = 0;
    unsigned char lo = 0x80, hi = 0xBF;  // allowed range of the second byte
    if (c >= 0xC2 && c <= 0xDF) need = 1;
    else if (c >= 0xE0 && c <= 0xEF) { need = 2; if (c == 0xE0) lo = 0xA0; if (c == 0xED) hi = 0x9F; }
    else if (c >= 0xF0 && c <= 0xF4) { need = 3; if (c == 0xF0) lo = 0x90; if (c == 0xF4) hi = 0x8F; }
    int have = 0;
    while (need > 0 && have < need && p + 1 + have < end) {
      const unsigned char n = p[1 + have];
      const bool ok = have == 0 ? (n >= lo && n <= hi) : ((n & 0xC0) == 0x80);
      if (!ok) break;
      ++have;
    }
    if (need > 0 && have == need) { out.append((const char*)p, need + 1); p += need + 1; }
    else { out += "\xEF\xBF\xBD"; p += 1 + have; }
  }
  out += '"';
}

// ------------------------------------------------------------------------------------- HTTP
struct Request {
  std::string method, target, path;  // path = percent-decoded, without query
  std::map<std::string, std::string> headers;  // lower-cased names
  bool keep_alive = true;
};

struct Conn {
  int fd;
  std::string buf;  // bytes read but not consumed
  size_t pos = 0;

  bool fill() {
    char tmp[65536];
    ssize_t n;
    do n = ::recv(fd, tmp, sizeof tmp, 0); while (n < 0 && errno == EINTR);
    if (n <= 0) return false;
    if (pos > 0 && pos == buf.size()) { buf.clear(); pos = 0; }
    buf.append(tmp, (size_t)n);
    return true;
  }
  // Reads up to and including "\r\n"; false on EOF / overlong line.
  bool line(std::string& out, size_t limit = 65536) {
    for (;;) {
      size_t e = buf.find("\r\n", pos);
      if (e != std::string::npos) { out.assign(buf, pos, e - pos); pos = e + 2; return true; }
      if (buf.size() - pos > limit) return false;
      if (!fill()) return false;
    }
  }
  // Hands out up to `want` body bytes that are already buffered or freshly received.
  ssize_t some(const char** data, size_t want) {
    if (pos == buf.size()) {
      buf.clear(); pos = 0;
      if (!fill()) return -1;
    }
    size_t n = std::min(want, buf.size() - pos);
    *data = buf.data() + pos;
    pos += n;
    return (ssize_t)n;
  }
  bool send_all(const char* p, size_t n) {
    while (n) {
      ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
      if (w < 0) { if (errno == EINTR) continue; return false; }
      p += w; n -= (size_t)w;
    }
    return true;
  }
  bool send_all(const std::string& s) { return send_all(s.data(), s.size()); }
};

std::string lower(std::string s) { for (auto& c : s) c = (char)tolower((unsigned char)c); return s; }

bool percent_decode(const std::string& in, std::string& out) {
  for (size_t i = 0; i < in.size(); ++i) {
    if (in[i] != '%') { out += in[i]; continue; }
    if (i + 2 >= in.size()) return false;
    auto hv = [](char c) -> int { if (c >= '0' && c <= '9') return c - '0'; c = (char)tolower(c); if (c >= 'a' && c <= 'f') return c - 'a' + 10; return -1; };
    int a = hv(in[i + 1]), b = hv(in[i + 2]);
    if (a < 0 || b < 0) return false;
    out += (char)(a * 16 + b);
    i += 2;
  }
  return true;
}

bool read_request(Conn& c, Request& r) {
  std::string l;
  do { if (!c.line(l)) return false; } while (l.empty());  // tolerate a stray CRLF between requests
  size_t a = l.find(' '), b = l.rfind(' ');
  if (a == std::string::npos || b == a) return false;
  r.method = l.substr(0, a);
  r.target = l.substr(a + 1, b - a - 1);
  const std::string version = l.substr(b + 1);
  r.keep_alive = version != "HTTP/1.0";
  size_t total = 0;
  for (;;) {
    if (!c.line(l)) return false;
    if (l.empty()) break;
    total += l.size();
    if (total > (1u << 20)) return false;
    size_t colon = l.find(':');
    if (colon == std::string::npos) continue;
    std::string v = l.substr(colon + 1);
    size_t s = v.find_first_not_of(" \t"), e = v.find_last_not_of(" \t");
    v = s == std::string::npos ? "" : v.substr(s, e - s + 1);
    r.headers[lower(l.substr(0, colon))] = v;
  }
  auto it = r.headers.find("connection");
  if (it != r.headers.end()) {
    std::string v = lower(it->second);
    if (v.find("close") != std::string::npos) r.keep_alive = false;
    else if (v.find("keep-alive") != std::string::npos) r.keep_alive = true;
  }
  std::string raw = r.target.substr(0, r.target.find('?'));
  r.path.clear();
  return percent_decode(raw, r.path);
}

// Streams the request body to `sink(data, n)` (false aborts).  Returns false on a protocol error
// or a dropped connection; `*sink_failed` tells the two apart from a sink refusal.
template <typename Sink>
bool read_body(Conn& c, const Request& r, Sink sink, bool* sink_failed) {
  *sink_failed = false;
  auto te = r.headers.find("transfer-encoding");
  if (te != r.headers.end() && lower(te->second).find("chunked") != std::string::npos) {
    for (;;) {
      std::string l;
      if (!c.line(l)) return false;
      unsigned long long n = strtoull(l.c_str(), nullptr, 16);
      if (n == 0) {
        do { if (!c.line(l)) return false; } while (!l.empty());  // trailers
        return true;
      }
      while (n) {
        const char* d;
        ssize_t got = c.some(&d, (size_t)std::min<unsigned long long>(n, 1u << 20));
        if (got < 0) return false;
        if (!*sink_failed && !sink(d, (size_t)got)) *sink_failed = true;
        n -= (unsigned long long)got;
      }
      if (!c.line(l)) return false;  // CRLF after the chunk
    }
  }
  unsigned long long n = 0;
  auto cl = r.headers.find("content-length");
  if (cl != r.headers.end()) n = strtoull(cl->second.c_str(), nullptr, 10);
  while (n) {
    const char* d;
    ssize_t got = c.some(&d, (size_t)std::min<unsigned long long>(n, 1u << 20));
    if (got < 0) return false;
    if (!*sink_failed && !sink(d, (size_t)got)) *sink_failed = true;
    n -= (unsigned long long)got;
  }
  return true;
}

bool respond(Conn& c, int status, const char* reason, const char* ctype, const std::string& body, bool keep_alive) {
  std::string h = "HTTP/1.1 " + std::to_string(status) + " " + reason + "\r\n";
  if (status != 204) {
    if (ctype) h += std::string("content-type: ") + ctype + "\r\n";
    h += "content-length: " + std::to_string(body.size()) + "\r\n";
  }
  h += keep_alive ? "connection: keep-alive\r\n\r\n" : "connection: close\r\n\r\n";
  if (status != 204) h += body;
  return c.send_all(h);
}

// /workspace/<rel> -> <workspace>/<rel>; refuses empty names and `..` components.
bool workspace_path(const std::string& url_path, std::string& out) {
  static const std::string prefix = "/workspace/";
  if (url_path.compare(0, prefix.size(), prefix) != 0) return false;
  std::string rel = url_path.substr(prefix.size());
  if (rel.empty() || rel.find('\0') != std::string::npos) return false;
  size_t i = 0;
  while (i <= rel.size()) {
    size_t j = rel.find('/', i);
    if (j == std::string::npos) j = rel.size();
    if (rel.compare(i, j - i, "..") == 0) return false;
    i = j + 1;
  }
  out = g_cfg.workspace + "/" + rel;
  return true;
}

bool mkdir_parents(const std::string& file_path) {
  for (size_t i = 1; i < file_path.size(); ++i) {
    if (file_path[i] != '/') continue;
    std::string d = file_path.substr(0, i);
    if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------- execute
struct Outcome { std::string out, err; int exit_code = 0; };

std::string read_file(const std::string& path) {
  std::string s;
  int fd = open(path.c_str(), O_RDONLY | O_CLOEXEC);
  if (fd < 0) return s;
  char tmp[65536];
  ssize_t n;
  while ((n = read(fd, tmp, sizeof tmp)) > 0) s.append(tmp, (size_t)n);
  close(fd);
  return s;
}

long long now_ms() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (long long)ts.tv_sec * 1000 + ts.tv_nsec / 1000000;
}

bool recv_exact(int fd, void* p, size_t n, long long deadline_ms, bool* timed_out) {
  char* c = (char*)p;
  *timed_out = false;
  while (n) {
    if (deadline_ms >= 0) {
      long long left = deadline_ms - now_ms();
      if (left <= 0) { *timed_out = true; return false; }
      struct pollfd pf = {fd, POLLIN, 0};
      int pr = poll(&pf, 1, (int)std::min<long long>(left, 1 << 30));
      if (pr == 0) { *timed_out = true; return false; }
      if (pr < 0) { if (errno == EINTR) continue; return false; }
    }
    ssize_t r = ::recv(fd, c, n, 0);
    if (r == 0) return false;
    if (r < 0) { if (errno == EINTR) continue; return false; }
    c += r; n -= (size_t)r;
  }
  return true;
}

static bool write_script(const std::string& script, const std::string& source, std::string& error);

// The zygote's child calls setsid(): its pid is the id of the script's process group.
static void kill_script_group(int32_t pid) {
  if (pid <= 1) return;  // never signal 0 / -1 on a garbled reply
  kill(-pid, SIGKILL);
  kill(pid, SIGKILL);
}

// Warm path: ask the zygote (zygote.py protocol, JSON flavour) to fork a child for the script.
bool run_via_zygote(const std::string& source, const std::map<std::string, std::string>& env, unsigned long long timeout_s,
                    const std::string& dir, const std::string& out_path, const std::string& err_path, Outcome& oc,
                    std::string& error) {
  const std::string script_path = dir + "/script.py";
  if (!write_script(script_path, source, error)) return false;
  int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) { error = "socket(): " + std::string(strerror(errno)); return false; }
  struct sockaddr_un sa;
  memset(&sa, 0, sizeof sa);
  sa.sun_family = AF_UNIX;
  if (g_cfg.zygote_socket.size() >= sizeof sa.sun_path) { close(fd); error = "zygote socket path too long"; return false; }
  strcpy(sa.sun_path, g_cfg.zygote_socket.c_str());
  if (connect(fd, (struct sockaddr*)&sa, sizeof sa) != 0) { error = "zygote connect: " + std::string(strerror(errno)); close(fd); return false; }
  std::string req = "{\"source_code\":";
  json_string(req, source);
  req += ",\"workspace\":"; json_string(req, g_cfg.workspace);
  req += ",\"stdout_path\":"; json_string(req, out_path);
  req += ",\"stderr_path\":"; json_string(req, err_path);
  req += ",\"script_path\":"; json_string(req, script_path);
  if (!g_cfg.broker_socket.empty()) { req += ",\"broker_socket\":"; json_string(req, g_cfg.broker_socket); }
  req += ",\"env\":{";
  bool first = true;
  for (auto& kv : env) {
    if (!first) req += ',';
    first = false;
    json_string(req, kv.first); req += ':'; json_string(req, kv.second);
  }
  req += "}}";
  uint32_t len = (uint32_t)req.size();
  std::string frame((const char*)&len, 4);
  frame += req;
  const char* p = frame.data();
  size_t n = frame.size();
  while (n) {
    ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w < 0) { if (errno == EINTR) continue; error = "zygote send failed"; close(fd); return false; }
    p += w; n -= (size_t)w;
  }
  int32_t pid = 0, code = 0;
  bool to = false;
  if (!recv_exact(fd, &pid, 4, now_ms() + 10000, &to)) { error = "zygote did not fork"; close(fd); return false; }
  const long long deadline = now_ms() + (long long)std::min<unsigned long long>(timeout_s, 1ull << 40) * 1000;
  if (!recv_exact(fd, &code, 4, deadline, &to)) {
    close(fd);
    if (to) {
      kill_script_group(pid);
      oc.out.clear(); oc.err = "Execution timed out"; oc.exit_code = -1;  // server.rs:169
      return true;
    }
    code = -1;  // died without reporting: killed by a signal (server.rs:165)
    kill_script_group(pid);
    oc.out = read_file(out_path); oc.err = read_file(err_path); oc.exit_code = -1;
    return true;
  }
  close(fd);
  kill_script_group(pid);  // whatever the script left running goes with the request (the reference deletes the pod)
  oc.out = read_file(out_path); oc.err = read_file(err_path);
  oc.exit_code = code < 0 ? -1 : code;
  return true;
}

// server.rs:124: the source goes into a file of the request's scratch directory.
static bool write_script(const std::string& script, const std::string& source, std::string& error) {
  int fd = open(script.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0600);
  if (fd < 0) { error = "cannot write script: " + std::string(strerror(errno)); return false; }
  const char* p = source.data();
  size_t n = source.size();
  while (n) {
    ssize_t w = write(fd, p, n);
    if (w < 0) { if (errno == EINTR) continue; close(fd); error = "cannot write script"; return false; }
    p += w; n -= (size_t)w;
  }
  close(fd);
  return true;
}

// Cold path: `$BCI_PYTHON script.py` as a child process, cwd = workspace.
bool run_via_exec(const std::string& source, const std::map<std::string, std::string>& env, unsigned long long timeout_s,
                  const std::string& dir, const std::string& out_path, const std::string& err_path, Outcome& oc,
                  std::string& error) {
  const std::string script = dir + "/script.py";
  if (!write_script(script, source, error)) return false;
  pid_t pid = fork();
  if (pid < 0) { error = "fork(): " + std::string(strerror(errno)); return false; }
  if (pid == 0) {
    setsid();
    if (chdir(g_cfg.workspace.c_str()) != 0) _exit(127);
    int o = open(out_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0600);
    int e = open(err_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0600);
    int z = open("/dev/null", O_RDONLY);
    if (o < 0 || e < 0 || z < 0) _exit(127);
    dup2(z, 0); dup2(o, 1); dup2(e, 2);
    for (auto& kv : env) setenv(kv.first.c_str(), kv.second.c_str(), 1);
    if (!g_cfg.broker_socket.empty()) setenv("BCI_BROKER_SOCKET", g_cfg.broker_socket.c_str(), 1);
    execlp(g_cfg.python.c_str(), g_cfg.python.c_str(), script.c_str(), (char*)nullptr);
    dprintf(2, "cannot start %s: %s\n", g_cfg.python.c_str(), strerror(errno));
    _exit(127);
  }
  const long long deadline = now_ms() + (long long)std::min<unsigned long long>(timeout_s, 1ull << 40) * 1000;
  int status = 0;
  long long nap_us = 200;
  for (;;) {
    pid_t w = waitpid(pid, &status, WNOHANG);
    if (w == pid) break;
    if (w < 0 && errno != EINTR) { error = "waitpid failed"; return false; }
    if (now_ms() >= deadline) {
      kill(-pid, SIGKILL);
      kill(pid, SIGKILL);
      waitpid(pid, &status, 0);
      oc.out.clear(); oc.err = "Execution timed out"; oc.exit_code = -1;
      return true;
    }
    usleep((useconds_t)nap_us);
    if (nap_us < 5000) nap_us *= 2;
  }
  kill(-pid, SIGKILL);   // whatever the script left running goes with the request (the reference deletes the pod)
  oc.out = read_file(out_path); oc.err = read_file(err_path);
  oc.exit_code = WIFEXITED(status) ? WEXITSTATUS(status) : -1;
  return true;
}

// What a change to a file cannot leave untouched.
struct FileSig {
  ino_t ino; off_t size; long long mtime_ns, ctime_ns;
  bool operator==(const FileSig& o) const { return ino == o.ino && size == o.size && mtime_ns == o.mtime_ns && ctime_ns == o.ctime_ns; }
};
static FileSig sig_of(const struct stat& st) {
  return FileSig{st.st_ino, st.st_size, (long long)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec,
                 (long long)st.st_ctim.tv_sec * 1000000000ll + st.st_ctim.tv_nsec};
}

void snapshot_workspace(std::map<std::string, FileSig>& snap) {
  DIR* d = opendir(g_cfg.workspace.c_str());
  if (!d) return;
  while (struct dirent* e = readdir(d)) {
    if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
    struct stat st;
    if (stat((g_cfg.workspace + "/" + e->d_name).c_str(), &st) == 0 && S_ISREG(st.st_mode)) snap[e->d_name] = sig_of(st);
  }
  closedir(d);
}

// server.rs:98-118: regular files directly in the workspace whose ctime is later than the start of
// the request.  The reference's script needs >= 80 ms to start, ours ~2 ms — inside one tick of the
// coarse clock the kernel stamps ctime with — so a file written right away can carry a ctime just
// BEFORE `since`.  Hence also: anything new or different from the snapshot taken at the start.
void changed_files(const struct timespec& since, const std::map<std::string, FileSig>& before, std::vector<std::string>& out) {
  DIR* d = opendir(g_cfg.workspace.c_str());
  if (!d) return;
  while (struct dirent* e = readdir(d)) {
    if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
    std::string p = g_cfg.workspace + "/" + e->d_name;
    struct stat st;
    if (stat(p.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;   // follows symlinks like Path::is_file
    const bool later = st.st_ctim.tv_sec > since.tv_sec || (st.st_ctim.tv_sec == since.tv_sec && st.st_ctim.tv_nsec > since.tv_nsec);
    auto it = before.find(e->d_name);
    if (later || it == before.end() || !(it->second == sig_of(st)))
      out.push_back((g_cfg.report_as.empty() ? g_cfg.workspace : g_cfg.report_as) + "/" + e->d_name);
  }
  closedir(d);
  std::sort(out.begin(), out.end());
}

void remove_tree_flat(const std::string& dir) {
  DIR* d = opendir(dir.c_str());
  if (d) {
    while (struct dirent* e = readdir(d)) {
      if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
      unlink((dir + "/" + e->d_name).c_str());
    }
    closedir(d);
  }
  rmdir(dir.c_str());
}

// Returns the HTTP status; `body` is the JSON result or the error text.
int handle_execute(const std::string& payload, std::string& body) {
  JsonParser jp{payload.data(), payload.data() + payload.size(), "", 0};
  JsonValue root;
  if (!jp.value(root)) { body = "Json deserialize error: " + jp.err; return 400; }
  jp.ws();
  if (jp.p != jp.end) { body = "Json deserialize error: trailing characters"; return 400; }
  if (root.kind != JsonValue::Object) { body = "Json deserialize error: expected an object"; return 400; }
  const JsonValue* src = root.get("source_code");
  if (!src || src->kind != JsonValue::String) { body = "Json deserialize error: missing field `source_code`"; return 400; }
  unsigned long long timeout_s = 60;  // server.rs:150
  if (const JsonValue* t = root.get("timeout")) {
    if (t->kind == JsonValue::Number && t->num_is_uint) timeout_s = t->u;
    else if (t->kind != JsonValue::Null) { body = "Json deserialize error: `timeout` must be a non-negative integer"; return 400; }
  }
  std::map<std::string, std::string> env;
  if (const JsonValue* e = root.get("env")) {
    if (e->kind == JsonValue::Object) {
      for (auto& kv : e->obj) {
        if (kv.second.kind != JsonValue::String) { body = "Json deserialize error: `env` values must be strings"; return 400; }
        env[kv.first] = kv.second.str;
      }
    } else if (e->kind != JsonValue::Null) { body = "Json deserialize error: `env` must be an object"; return 400; }
  }

  struct timespec start;
  clock_gettime(CLOCK_REALTIME, &start);
  std::map<std::string, FileSig> before;
  snapshot_workspace(before);
  std::string tmpl = g_cfg.scratch + "/bci-exec-XXXXXX";
  std::vector<char> tb(tmpl.begin(), tmpl.end());
  tb.push_back('\0');
  if (!mkdtemp(tb.data())) { body = "cannot create a scratch directory: " + std::string(strerror(errno)); return 500; }
  const std::string dir = tb.data();
  const std::string out_path = dir + "/stdout", err_path = dir + "/stderr";
  Outcome oc;
  std::string error;
  bool ok = g_cfg.zygote_socket.empty()
                ? run_via_exec(src->str, env, timeout_s, dir, out_path, err_path, oc, error)
                : run_via_zygote(src->str, env, timeout_s, dir, out_path, err_path, oc, error);
  remove_tree_flat(dir);
  if (!ok) { body = error; return 500; }
  std::vector<std::string> files;
  changed_files(start, before, files);
  body = "{\"stdout\":";
  json_string(body, oc.out);
  body += ",\"stderr\":";
  json_string(body, oc.err);
  body += ",\"exit_code\":" + std::to_string(oc.exit_code) + ",\"files\":[";
  for (size_t i = 0; i < files.size(); ++i) {
    if (i) body += ',';
    json_string(body, files[i]);
  }
  body += "]}";
  return 200;
}

// ------------------------------------------------------------------------------------- routes
void serve_connection(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
  Conn c{fd, std::string(), 0};
  for (;;) {
    Request r;
    if (!read_request(c, r)) break;
    auto ex = r.headers.find("expect");
    if (ex != r.headers.end() && lower(ex->second) == "100-continue") c.send_all("HTTP/1.1 100 Continue\r\n\r\n");
    bool sink_failed = false;
    bool alive = true;
    if (r.method == "PUT" && r.path.compare(0, 11, "/workspace/") == 0) {
      std::string target;
      int status = 204;
      std::string msg;
      int out = -1;
      if (!workspace_path(r.path, target)) { status = 400; msg = "bad workspace path"; }
      else if (!mkdir_parents(target)) { status = 500; msg = std::string("cannot create directories: ") + strerror(errno); }
      else {
        out = open(target.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
        if (out < 0) { status = 500; msg = std::string("cannot open file: ") + strerror(errno); }
      }
      alive = read_body(c, r, [&](const char* d, size_t n) {
        if (out < 0) return true;  // drain
        while (n) {
          ssize_t w = write(out, d, n);
          if (w < 0) { if (errno == EINTR) continue; return false; }
          d += w; n -= (size_t)w;
        }
        return true;
      }, &sink_failed);
      if (out >= 0) close(out);
      if (!alive) break;
      if (sink_failed) { status = 500; msg = "write failed"; }
      if (!respond(c, status, status == 204 ? "No Content" : (status == 400 ? "Bad Request" : "Internal Server Error"), "text/plain", msg, r.keep_alive)) break;
    } else if ((r.method == "GET" || r.method == "HEAD") && r.path.compare(0, 11, "/workspace/") == 0) {
      alive = read_body(c, r, [](const char*, size_t) { return true; }, &sink_failed);
      if (!alive) break;
      std::string target;
      if (!workspace_path(r.path, target)) { if (!respond(c, 400, "Bad Request", "text/plain", "bad workspace path", r.keep_alive)) break; goto next; }
      {
        int in = open(target.c_str(), O_RDONLY | O_CLOEXEC);
        struct stat st;
        if (in < 0 || fstat(in, &st) != 0 || !S_ISREG(st.st_mode)) {
          int status = (in < 0 && errno == EACCES) ? 403 : 404;
          if (in >= 0) close(in);
          if (!respond(c, status, status == 403 ? "Forbidden" : "Not Found", "text/plain", "", r.keep_alive)) break;
          goto next;
        }
        std::string h = "HTTP/1.1 200 OK\r\ncontent-type: application/octet-stream\r\ncontent-length: " +
                        std::to_string((unsigned long long)st.st_size) + "\r\n" +
                        (r.keep_alive ? "connection: keep-alive\r\n\r\n" : "connection: close\r\n\r\n");
        bool ok = c.send_all(h);
        if (ok && r.method == "GET") {
          std::vector<char> tmp(1 << 20);
          unsigned long long left = (unsigned long long)st.st_size;
          while (ok && left) {
            ssize_t n = read(in, tmp.data(), (size_t)std::min<unsigned long long>(left, tmp.size()));
            if (n <= 0) { ok = false; break; }  // file shrank: the length promised cannot be met
            ok = c.send_all(tmp.data(), (size_t)n);
            left -= (unsigned long long)n;
          }
        }
        close(in);
        if (!ok) break;
      }
    } else if (r.method == "POST" && r.path == "/execute") {
      std::string payload;
      bool too_big = false;
      alive = read_body(c, r, [&](const char* d, size_t n) {
        if (payload.size() + n > (256u << 20)) { too_big = true; return false; }
        payload.append(d, n);
        return true;
      }, &sink_failed);
      if (!alive) break;
      if (too_big) { if (!respond(c, 413, "Payload Too Large", "text/plain", "request body too large", r.keep_alive)) break; goto next; }
      std::string body;
      int status = handle_execute(payload, body);
      const char* reason = status == 200 ? "OK" : (status == 400 ? "Bad Request" : "Internal Server Error");
      if (!respond(c, status, reason, status == 200 ? "application/json" : "text/plain", body, r.keep_alive)) break;
    } else {
      alive = read_body(c, r, [](const char*, size_t) { return true; }, &sink_failed);
      if (!alive) break;
      bool known = r.path == "/execute" || r.path.compare(0, 11, "/workspace/") == 0;
      if (!respond(c, known ? 405 : 404, known ? "Method Not Allowed" : "Not Found", "text/plain", "", r.keep_alive)) break;
    }
  next:
    if (!r.keep_alive) break;
  }
  shutdown(fd, SHUT_RDWR);
  close(fd);
}

std::atomic<int> g_listen_fd{-1};
void on_term(int) {
  int fd = g_listen_fd.exchange(-1);
  if (fd >= 0) close(fd);
  _exit(0);
}

}  // namespace

int main() {
  if (const char* v = getenv("APP_LISTEN_ADDR")) {
    std::string a = v;
    size_t colon = a.rfind(':');
    if (colon == std::string::npos) { fprintf(stderr, "APP_LISTEN_ADDR must be host:port\n"); return 2; }
    g_cfg.host = a.substr(0, colon);
    g_cfg.port = atoi(a.c_str() + colon + 1);
  }
  if (const char* v = getenv("APP_WORKSPACE")) g_cfg.workspace = v;
  while (g_cfg.workspace.size() > 1 && g_cfg.workspace.back() == '/') g_cfg.workspace.pop_back();
  if (const char* v = getenv("BCI_ZYGOTE_SOCKET")) g_cfg.zygote_socket = v;
  if (const char* v = getenv("BCI_PYTHON")) g_cfg.python = v;
  if (const char* v = getenv("BCI_BROKER_SOCKET")) g_cfg.broker_socket = v;
  if (const char* v = getenv("BCI_WORKSPACE_REPORT_AS")) g_cfg.report_as = v;
  while (g_cfg.report_as.size() > 1 && g_cfg.report_as.back() == '/') g_cfg.report_as.pop_back();
  if (const char* v = getenv("BCI_SCRATCH")) g_cfg.scratch = v;
  else { struct stat st; if (stat("/dev/shm", &st) == 0 && S_ISDIR(st.st_mode) && access("/dev/shm", W_OK) == 0) g_cfg.scratch = "/dev/shm"; }
  signal(SIGPIPE, SIG_IGN);
  signal(SIGTERM, on_term);
  signal(SIGINT, on_term);

  int fd = socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) { perror("socket"); return 1; }
  int one = 1;
  setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  struct sockaddr_in sa;
  memset(&sa, 0, sizeof sa);
  sa.sin_family = AF_INET;
  sa.sin_port = htons((uint16_t)g_cfg.port);
  if (inet_pton(AF_INET, g_cfg.host.c_str(), &sa.sin_addr) != 1) { fprintf(stderr, "bad listen host %s\n", g_cfg.host.c_str()); return 2; }
  if (bind(fd, (struct sockaddr*)&sa, sizeof sa) != 0) { perror("bind"); return 1; }
  if (listen(fd, 1024) != 0) { perror("listen"); return 1; }
  socklen_t sl = sizeof sa;
  getsockname(fd, (struct sockaddr*)&sa, &sl);
  g_listen_fd = fd;
  printf("listening on %s:%d workspace=%s mode=%s\n", g_cfg.host.c_str(), (int)ntohs(sa.sin_port), g_cfg.workspace.c_str(),
         g_cfg.zygote_socket.empty() ? "exec" : "zygote");
  fflush(stdout);
  for (;;) {
    int cfd = accept4(fd, nullptr, nullptr, SOCK_CLOEXEC);
    if (cfd < 0) {
      if (errno == EINTR || errno == ECONNABORTED) continue;
      if (errno == EMFILE || errno == ENFILE) { usleep(1000); continue; }
      break;
    }
    std::thread(serve_connection, cfd).detach();
  }
  return 0;
}
