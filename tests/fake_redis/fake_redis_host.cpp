// fake_redis_host — TEST INFRASTRUCTURE: a tiny stand-in for redis-server's module host.
//
// No redis-server / redismodule.h exists in this image, so the Redis module (redis_hnsw_b200/csrc/redis/) is exercised
// against this host: it dlopen()s the module, hands RedisModule_OnLoad a context whose first word is the GetApi
// resolver (the module ABI convention), provides the function table the module asks for (keyspace of module-typed
// values, reply builder, RDB-style typed IO) and runs a command script:
//
//     fake_redis_host <module.so> < script
//
// Each script line is one command (whitespace-separated words).  Meta commands:
//     #SAVE <file>     rdb_save every module-typed key into <file> (in this process, like the SAVE command)
//     #BGSAVE <file>   like BGSAVE: fork(); the CHILD fires the persistence server event (RDB_START — redis-server calls
//                      startSaving() from rdbSave(), i.e. inside the forked child; only SAVE / SYNC_RDB_START run in the
//                      parent) and rdb_saves every key; it must not touch the parent's CUDA context.  The parent waits
//                      and fires the end event.  FAKE_REDIS_NO_EVENTS=1 emulates a host without server events.
//     #LOAD <file>     rdb_load the keys of <file> into the (empty) keyspace
//     #KEYS            reply: sorted [key, type-name] pairs
//     #INFO            reply: module name/version, registered commands (name, flags, key spec) and data types
//     #FLUSHALL        delete every key (free callbacks run)
// Every command prints ONE line of JSON: integers as JSON ints, doubles as JSON numbers that always carry a '.', 'e',
// "inf" or "nan" marker ({"double": "..."} for non-finite), bulk strings as JSON strings, simple strings as
// {"status": ...}, errors as {"error": ...}, null as null, arrays as arrays.
#include <dlfcn.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

struct RedisModuleString {
  std::string s;
};

struct TypeMethods {
  uint64_t version;
  void* (*rdb_load)(void* io, int encver);
  void (*rdb_save)(void* io, void* value);
  void* aof_rewrite;
  void* mem_usage;
  void* digest;
  void (*free)(void* value);
};

struct RedisModuleType {
  std::string name;
  int encver;
  TypeMethods m;
};

struct Entry {
  RedisModuleType* type = nullptr;
  void* value = nullptr;
};

struct Command {
  int (*fn)(void* ctx, RedisModuleString** argv, int argc);
  std::string flags;
  int first, last, step;
};

struct ServerEvent {
  uint64_t id, dataver;
};
typedef void (*EventCallback)(void* ctx, ServerEvent eid, uint64_t subevent, void* data);

struct Host {
  std::vector<std::pair<uint64_t, EventCallback>> subscribers;
  std::string module_name;
  int module_ver = 0, api_ver = 0;
  std::map<std::string, Command> commands;
  std::map<std::string, RedisModuleType*> types;
  std::map<std::string, Entry> keys;
} H;

struct RedisModuleKey {
  std::string name;
  int mode;
};

// reply tree
struct Reply {
  enum Kind { kInt, kDouble, kBulk, kStatus, kError, kNull, kArray } kind = kNull;
  long long i = 0;
  double d = 0;
  std::string s;
  std::vector<Reply> items;
  long want = 0;  // announced array length
};

struct Ctx {
  void* get_api;  // MUST be the first word: RedisModule_Init reads ((void**)ctx)[0]
  Reply root;
  bool have_root = false;
  std::vector<Reply*> open;  // arrays still expecting items
  std::vector<RedisModuleString*> auto_strings;
  bool auto_memory = false;
};

static void add_reply(Ctx* c, Reply r) {
  Reply* slot;
  if (c->open.empty()) {
    c->root = std::move(r);
    c->have_root = true;
    slot = &c->root;
  } else {
    Reply* parent = c->open.back();
    parent->items.push_back(std::move(r));
    slot = &parent->items.back();
    // completing containers is handled below through indices, because push_back may move earlier siblings
  }
  if (slot->kind == Reply::kArray && slot->want > 0) {
    c->open.push_back(slot);
    return;
  }
  // close every array that just received its last item
  while (!c->open.empty() && (long)c->open.back()->items.size() == c->open.back()->want) c->open.pop_back();
}

// ---- RDB-style typed IO
struct IO {
  std::string buf;
  size_t pos = 0;
  bool error = false;
  void put(char tag, const void* p, size_t n) {
    buf.push_back(tag);
    buf.append((const char*)p, n);
  }
  bool get(char tag, void* p, size_t n) {
    if (pos + 1 + n > buf.size() || buf[pos] != tag) {
      error = true;
      std::memset(p, 0, n);
      return false;
    }
    std::memcpy(p, buf.data() + pos + 1, n);
    pos += 1 + n;
    return true;
  }
};

extern "C" {

static void* api_Alloc(size_t n) { return std::malloc(n); }
static void api_Free(void* p) { std::free(p); }

static int api_CreateCommand(void*, const char* name, int (*fn)(void*, RedisModuleString**, int), const char* flags, int first,
                             int last, int step) {
  if (H.commands.count(name)) return 1;
  H.commands[name] = Command{fn, flags ? flags : "", first, last, step};
  return 0;
}
static void api_SetModuleAttribs(void*, const char* name, int ver, int apiver) {
  H.module_name = name;
  H.module_ver = ver;
  H.api_ver = apiver;
}
static int api_IsModuleNameBusy(const char*) { return 0; }
static int api_WrongArity(void* ctx) {
  Reply r;
  r.kind = Reply::kError;
  r.s = "ERR wrong number of arguments";
  add_reply((Ctx*)ctx, r);
  return 0;
}
static void api_AutoMemory(void* ctx) { ((Ctx*)ctx)->auto_memory = true; }

static RedisModuleType* api_CreateDataType(void*, const char* name, int encver, TypeMethods* m) {
  if (std::strlen(name) != 9 || H.types.count(name)) return nullptr;  // redis insists on 9-character type names
  RedisModuleType* t = new RedisModuleType{name, encver, TypeMethods{}};
  t->m.version = m->version;
  t->m.rdb_load = m->rdb_load;
  t->m.rdb_save = m->rdb_save;
  t->m.aof_rewrite = m->aof_rewrite;
  t->m.mem_usage = m->mem_usage;
  t->m.digest = m->digest;
  t->m.free = m->free;
  H.types[name] = t;
  return t;
}

static void* api_OpenKey(void*, RedisModuleString* name, int mode) {
  if (!(mode & 2) && !H.keys.count(name->s)) return nullptr;  // read-only open of a missing key
  return new RedisModuleKey{name->s, mode};
}
static void api_CloseKey(RedisModuleKey* k) { delete k; }
static int api_KeyType(RedisModuleKey* k) {
  if (!k || !H.keys.count(k->name)) return 0;
  return 6;
}
static void drop_key(const std::string& name) {
  auto it = H.keys.find(name);
  if (it == H.keys.end()) return;
  Entry e = it->second;
  H.keys.erase(it);
  if (e.type && e.type->m.free && e.value) e.type->m.free(e.value);
}
static int api_DeleteKey(RedisModuleKey* k) {
  if (!k || !(k->mode & 2)) return 1;
  drop_key(k->name);
  return 0;
}
static RedisModuleType* api_ModuleTypeGetType(RedisModuleKey* k) {
  if (!k) return nullptr;
  auto it = H.keys.find(k->name);
  return it == H.keys.end() ? nullptr : it->second.type;
}
static void* api_ModuleTypeGetValue(RedisModuleKey* k) {
  if (!k) return nullptr;
  auto it = H.keys.find(k->name);
  return it == H.keys.end() ? nullptr : it->second.value;
}
static int api_ModuleTypeSetValue(RedisModuleKey* k, RedisModuleType* t, void* value) {
  if (!k || !(k->mode & 2)) return 1;
  drop_key(k->name);
  H.keys[k->name] = Entry{t, value};
  return 0;
}

static RedisModuleString* api_CreateString(void* ctx, const char* p, size_t n) {
  RedisModuleString* s = new RedisModuleString{std::string(p, n)};
  if (ctx && ((Ctx*)ctx)->auto_memory) ((Ctx*)ctx)->auto_strings.push_back(s);
  return s;
}
static void api_FreeString(void* ctx, RedisModuleString* s) {
  if (ctx) {
    auto& v = ((Ctx*)ctx)->auto_strings;
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i] == s) {
        v.erase(v.begin() + i);
        break;
      }
  }
  delete s;
}
static const char* api_StringPtrLen(const RedisModuleString* s, size_t* len) {
  if (len) *len = s->s.size();
  return s->s.data();
}

static int api_ReplyWithError(void* ctx, const char* e) {
  Reply r;
  r.kind = Reply::kError;
  r.s = e;
  add_reply((Ctx*)ctx, r);
  return 0;
}
static int api_ReplyWithSimpleString(void* ctx, const char* m) {
  Reply r;
  r.kind = Reply::kStatus;
  r.s = m;
  add_reply((Ctx*)ctx, r);
  return 0;
}
static int api_ReplyWithLongLong(void* ctx, long long v) {
  Reply r;
  r.kind = Reply::kInt;
  r.i = v;
  add_reply((Ctx*)ctx, r);
  return 0;
}
static int api_ReplyWithDouble(void* ctx, double v) {
  Reply r;
  r.kind = Reply::kDouble;
  r.d = v;
  add_reply((Ctx*)ctx, r);
  return 0;
}
static int api_ReplyWithArray(void* ctx, long len) {
  Reply r;
  r.kind = Reply::kArray;
  r.want = len;
  add_reply((Ctx*)ctx, r);
  return 0;
}
static int api_ReplyWithStringBuffer(void* ctx, const char* p, size_t n) {
  Reply r;
  r.kind = Reply::kBulk;
  r.s.assign(p, n);
  add_reply((Ctx*)ctx, r);
  return 0;
}
static int api_ReplyWithNull(void* ctx) {
  add_reply((Ctx*)ctx, Reply());
  return 0;
}

static void api_SaveUnsigned(IO* io, uint64_t v) { io->put('U', &v, 8); }
static uint64_t api_LoadUnsigned(IO* io) {
  uint64_t v;
  io->get('U', &v, 8);
  return v;
}
static void api_SaveDouble(IO* io, double v) { io->put('D', &v, 8); }
static double api_LoadDouble(IO* io) {
  double v;
  io->get('D', &v, 8);
  return v;
}
static void api_SaveFloat(IO* io, float v) { io->put('F', &v, 4); }
static float api_LoadFloat(IO* io) {
  float v;
  io->get('F', &v, 4);
  return v;
}
static void api_SaveStringBuffer(IO* io, const char* p, size_t n) {
  uint64_t len = n;
  io->put('S', &len, 8);
  io->buf.append(p, n);
}
static char* api_LoadStringBuffer(IO* io, size_t* lenptr) {
  uint64_t len = 0;
  if (!io->get('S', &len, 8) || io->pos + len > io->buf.size()) {
    io->error = true;
    if (lenptr) *lenptr = 0;
    return nullptr;
  }
  char* out = (char*)std::malloc(len + 1);
  std::memcpy(out, io->buf.data() + io->pos, len);
  out[len] = 0;
  io->pos += len;
  if (lenptr) *lenptr = len;
  return out;
}

static int api_SubscribeToServerEvent(void*, ServerEvent ev, EventCallback cb) {
  H.subscribers.emplace_back(ev.id, cb);
  return 0;
}

static int get_api(const char* name, void* target) {
  if (std::string(name) == "RedisModule_SubscribeToServerEvent") {
    if (std::getenv("FAKE_REDIS_NO_EVENTS")) return 1;
    *(void**)target = (void*)api_SubscribeToServerEvent;
    return 0;
  }
  static const std::map<std::string, void*> table = {
#define E(n) {"RedisModule_" #n, (void*)api_##n}
      E(Alloc), E(Free), E(CreateCommand), E(SetModuleAttribs), E(IsModuleNameBusy), E(WrongArity), E(AutoMemory),
      E(CreateDataType), E(OpenKey), E(CloseKey), E(KeyType), E(DeleteKey), E(ModuleTypeGetType), E(ModuleTypeGetValue),
      E(ModuleTypeSetValue), E(CreateString), E(FreeString), E(StringPtrLen), E(ReplyWithError), E(ReplyWithSimpleString),
      E(ReplyWithLongLong), E(ReplyWithDouble), E(ReplyWithArray), E(ReplyWithStringBuffer), E(ReplyWithNull),
      E(SaveUnsigned), E(LoadUnsigned), E(SaveDouble), E(LoadDouble), E(SaveFloat), E(LoadFloat), E(SaveStringBuffer),
      E(LoadStringBuffer),
#undef E
  };
  auto it = table.find(name);
  if (it == table.end()) return 1;
  *(void**)target = it->second;
  return 0;
}

}  // extern "C"

// ---- JSON output
static void json_str(std::ostream& o, const std::string& s) {
  o << '"';
  for (unsigned char c : s) {
    if (c == '"') o << "\\\"";
    else if (c == '\\') o << "\\\\";
    else if (c == '\n') o << "\\n";
    else if (c < 0x20) {
      char b[8];
      std::snprintf(b, sizeof b, "\\u%04x", c);
      o << b;
    } else o << c;
  }
  o << '"';
}

static void json(std::ostream& o, const Reply& r) {
  switch (r.kind) {
    case Reply::kInt: o << r.i; break;
    case Reply::kDouble: {
      if (!std::isfinite(r.d)) {
        o << "{\"double\": \"" << (std::isnan(r.d) ? "nan" : (r.d > 0 ? "inf" : "-inf")) << "\"}";
        break;
      }
      char b[64];
      std::snprintf(b, sizeof b, "%.17g", r.d);
      std::string s = b;
      if (s.find_first_of(".e") == std::string::npos) s += ".0";
      o << s;
      break;
    }
    case Reply::kBulk: json_str(o, r.s); break;
    case Reply::kStatus: o << "{\"status\": ", json_str(o, r.s), o << "}"; break;
    case Reply::kError: o << "{\"error\": ", json_str(o, r.s), o << "}"; break;
    case Reply::kNull: o << "null"; break;
    case Reply::kArray:
      o << "[";
      for (size_t i = 0; i < r.items.size(); ++i) {
        if (i) o << ", ";
        json(o, r.items[i]);
      }
      o << "]";
      break;
  }
}

static std::string lower(std::string s) {
  for (char& c : s) c = (char)std::tolower((unsigned char)c);
  return s;
}

static void put_u64(std::string& b, uint64_t v) { b.append((const char*)&v, 8); }
static bool take_u64(const std::string& b, size_t& pos, uint64_t* v) {
  if (pos + 8 > b.size()) return false;
  std::memcpy(v, b.data() + pos, 8);
  pos += 8;
  return true;
}
static void put_s(std::string& b, const std::string& s) {
  put_u64(b, s.size());
  b += s;
}
static bool take_s(const std::string& b, size_t& pos, std::string* s) {
  uint64_t n;
  if (!take_u64(b, pos, &n) || pos + n > b.size()) return false;
  s->assign(b, pos, n);
  pos += n;
  return true;
}

static long long save_all(const std::string& path) {
  std::string out;
  uint64_t n = 0;
  for (auto& kv : H.keys) n += kv.second.type != nullptr;
  put_u64(out, n);
  for (auto& kv : H.keys) {
    if (!kv.second.type) continue;
    IO io;
    kv.second.type->m.rdb_save(&io, kv.second.value);
    put_s(out, kv.first);
    put_s(out, kv.second.type->name);
    put_u64(out, (uint64_t)kv.second.type->encver);
    put_s(out, io.buf);
  }
  std::ofstream f(path, std::ios::binary);
  f.write(out.data(), (std::streamsize)out.size());
  return f.good() ? (long long)n : -1;
}

static void fire_event(uint64_t id, uint64_t subevent) {
  Ctx c{};
  c.get_api = (void*)get_api;
  for (auto& sub : H.subscribers)
    if (sub.first == id) sub.second(&c, ServerEvent{id, 1}, subevent, nullptr);
}

static Reply meta(const std::vector<std::string>& w) {
  Reply r;
  if (w[0] == "#SAVE" && w.size() == 2) {
    r.kind = Reply::kInt;
    r.i = save_all(w[1]);
  } else if (w[0] == "#BGSAVE" && w.size() == 2) {
    std::fflush(stdout);
    pid_t pid = fork();
    if (pid == 0) {
      fire_event(1 /* REDISMODULE_EVENT_PERSISTENCE */, 0 /* RDB_START: fired by rdbSave() in the child */);
      long long n = save_all(w[1]);
      _exit(n >= 0 ? 0 : 1);  // no atexit handlers, no static destructors: like redis' child
    }
    int status = 0;
    waitpid(pid, &status, 0);
    fire_event(1, 4 /* ENDED on Redis >= 7.0 */);
    if (pid < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) {
      r.kind = Reply::kError;
      r.s = "ERR background save failed";
    } else {
      r.kind = Reply::kStatus;
      r.s = "Background saving done";
    }
  } else if (w[0] == "#LOAD" && w.size() == 2) {
    std::ifstream f(w[1], std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string in = ss.str();
    size_t pos = 0;
    uint64_t n = 0;
    bool ok = take_u64(in, pos, &n);
    long long loaded = 0;
    for (uint64_t i = 0; ok && i < n; ++i) {
      std::string key, type, payload;
      uint64_t encver = 0;
      ok = take_s(in, pos, &key) && take_s(in, pos, &type) && take_u64(in, pos, &encver) && take_s(in, pos, &payload);
      if (!ok) break;
      auto it = H.types.find(type);
      if (it == H.types.end()) {
        ok = false;
        break;
      }
      IO io;
      io.buf = payload;
      void* v = it->second->m.rdb_load(&io, (int)encver);
      if (!v || io.error || io.pos != io.buf.size()) {
        ok = false;
        break;
      }
      drop_key(key);
      H.keys[key] = Entry{it->second, v};
      ++loaded;
    }
    if (!ok) {
      r.kind = Reply::kError;
      r.s = "ERR rdb load failed";
    } else {
      r.kind = Reply::kInt;
      r.i = loaded;
    }
  } else if (w[0] == "#KEYS") {
    r.kind = Reply::kArray;
    for (auto& kv : H.keys) {
      Reply pair;
      pair.kind = Reply::kArray;
      Reply a, b;
      a.kind = b.kind = Reply::kBulk;
      a.s = kv.first;
      b.s = kv.second.type ? kv.second.type->name : "?";
      pair.items = {a, b};
      r.items.push_back(pair);
    }
  } else if (w[0] == "#INFO") {
    r.kind = Reply::kArray;
    Reply nm, ver, cmds, types;
    nm.kind = Reply::kBulk;
    nm.s = H.module_name;
    ver.kind = Reply::kInt;
    ver.i = H.module_ver;
    cmds.kind = types.kind = Reply::kArray;
    for (auto& c : H.commands) {
      Reply one;
      one.kind = Reply::kArray;
      Reply a, b, k1, k2, k3;
      a.kind = b.kind = Reply::kBulk;
      a.s = c.first;
      b.s = c.second.flags;
      k1.kind = k2.kind = k3.kind = Reply::kInt;
      k1.i = c.second.first, k2.i = c.second.last, k3.i = c.second.step;
      one.items = {a, b, k1, k2, k3};
      cmds.items.push_back(one);
    }
    for (auto& t : H.types) {
      Reply one;
      one.kind = Reply::kArray;
      Reply a, b;
      a.kind = Reply::kBulk;
      a.s = t.first;
      b.kind = Reply::kInt;
      b.i = t.second->encver;
      one.items = {a, b};
      types.items.push_back(one);
    }
    r.items = {nm, ver, cmds, types};
  } else if (w[0] == "#FLUSHALL") {
    while (!H.keys.empty()) drop_key(H.keys.begin()->first);
    r.kind = Reply::kStatus;
    r.s = "OK";
  } else {
    r.kind = Reply::kError;
    r.s = "ERR unknown meta command";
  }
  return r;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: fake_redis_host <module.so> < script\n");
    return 2;
  }
  void* so = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!so) {
    std::fprintf(stderr, "dlopen failed: %s\n", dlerror());
    return 3;
  }
  typedef int (*OnLoad)(void*, RedisModuleString**, int);
  OnLoad onload = (OnLoad)dlsym(so, "RedisModule_OnLoad");
  if (!onload) {
    std::fprintf(stderr, "RedisModule_OnLoad not exported\n");
    return 4;
  }
  {
    Ctx c{};
    c.get_api = (void*)get_api;
    if (onload(&c, nullptr, 0) != 0) {
      std::fprintf(stderr, "RedisModule_OnLoad failed\n");
      return 5;
    }
  }
  const bool timing = std::getenv("FAKE_REDIS_TIMING") != nullptr;
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream ss(line);
    std::vector<std::string> w;
    for (std::string t; ss >> t;) w.push_back(t);
    if (w.empty()) continue;
    Reply out;
    if (w[0][0] == '#') {
      out = meta(w);
    } else {
      auto it = H.commands.find(lower(w[0]));
      if (it == H.commands.end()) {
        out.kind = Reply::kError;
        out.s = "ERR unknown command '" + w[0] + "'";
      } else {
        Ctx c{};
        c.get_api = (void*)get_api;
        std::vector<RedisModuleString*> args;
        for (auto& t : w) args.push_back(new RedisModuleString{t});
        const auto t0 = std::chrono::steady_clock::now();
        it->second.fn(&c, args.data(), (int)args.size());
        if (timing)  // FAKE_REDIS_TIMING=1: wall time of the command handler alone (no script parsing, no reply printing)
          std::fprintf(stderr, "T %s %.1f\n", it->first.c_str(),
                       std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
        for (auto* a : args) delete a;
        for (auto* a : c.auto_strings) delete a;
        if (!c.have_root || !c.open.empty()) {
          out.kind = Reply::kError;
          out.s = "ERR module produced an incomplete reply";
        } else {
          out = c.root;
        }
      }
    }
    json(std::cout, out);
    std::cout << "\n" << std::flush;
  }
  while (!H.keys.empty()) drop_key(H.keys.begin()->first);  // free callbacks run like on FLUSHALL / shutdown
  return 0;
}
