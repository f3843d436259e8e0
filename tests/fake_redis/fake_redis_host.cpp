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
// Module values are serialised with redis-server's own RDB encoding for module types (rdb.c, RDB_TYPE_MODULE_2), so a
// payload written here is byte for byte what `hnsw.{index}` / `hnsw.{index}.{node}` hold inside a real dump.rdb — and a
// payload derived by hand from the reference's rdb_save callbacks (tests/golden/make_rdb_fixture.py) loads here:
//   value   := len(module id) item* len(0 = RDB_MODULE_OPCODE_EOF)
//   item    := len(2 = UINT) len(v) | len(3 = FLOAT) 4 bytes LE | len(4 = DOUBLE) 8 bytes LE | len(5 = STRING) string
//   len(v)  := 00vvvvvv | 01vvvvvv vvvvvvvv | 0x80 u32 big-endian | 0x81 u64 big-endian                    (rdbSaveLen)
//   string  := len(n) n bytes | 0xC0 int8 | 0xC1 int16 LE | 0xC2 int32 LE | 0xC3 len(clen) len(ulen) LZF bytes
//              (rdbSaveRawString: this host always WRITES the first form — what redis-server writes for strings that do not
//              look like integers and are <= 20 bytes, or with `rdbcompression no` — and READS all of them)
//   module id = 9 name characters, 6 bits each (A-Za-z0-9-_), then 10 bits of encoding version         (moduleTypeEncodeId)
enum { kOpEof = 0, kOpSint = 1, kOpUint = 2, kOpFloat = 3, kOpDouble = 4, kOpString = 5 };

static void rdb_put_len(std::string& b, uint64_t v) {
  if (v < (1u << 6)) {
    b.push_back((char)v);
  } else if (v < (1u << 14)) {
    b.push_back((char)(0x40 | (v >> 8)));
    b.push_back((char)(v & 0xFF));
  } else if (v <= 0xFFFFFFFFull) {
    b.push_back((char)0x80);
    for (int i = 3; i >= 0; --i) b.push_back((char)((v >> (8 * i)) & 0xFF));
  } else {
    b.push_back((char)0x81);
    for (int i = 7; i >= 0; --i) b.push_back((char)((v >> (8 * i)) & 0xFF));
  }
}

// returns false at the end of the buffer or on a special ("encoded") length, whose kind is stored in *enc
static bool rdb_get_len(const std::string& b, size_t& pos, uint64_t* v, int* enc = nullptr) {
  if (enc) *enc = -1;
  if (pos >= b.size()) return false;
  const unsigned char c = (unsigned char)b[pos++];
  const int type = c >> 6;
  if (type == 0) {
    *v = c & 0x3F;
  } else if (type == 1) {
    if (pos >= b.size()) return false;
    *v = ((uint64_t)(c & 0x3F) << 8) | (unsigned char)b[pos++];
  } else if (c == 0x80 || c == 0x81) {
    const int n = c == 0x80 ? 4 : 8;
    if (pos + n > b.size()) return false;
    *v = 0;
    for (int i = 0; i < n; ++i) *v = (*v << 8) | (unsigned char)b[pos++];
  } else {
    if (enc) *enc = c & 0x3F;  // 11xxxxxx: specially encoded string
    return false;
  }
  return true;
}

static const char* kModuleIdChars = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";
static uint64_t module_type_id(const std::string& name, int encver) {
  uint64_t id = 0;
  for (int i = 0; i < 9; ++i) {
    const char* p = std::strchr(kModuleIdChars, i < (int)name.size() ? name[i] : 'A');
    id = (id << 6) | (uint64_t)(p ? p - kModuleIdChars : 0);
  }
  return (id << 10) | (uint64_t)(encver & 1023);
}

// liblzf decompression (the format rdbSaveLzfStringObject writes)
static bool lzf_decompress(const unsigned char* in, size_t in_len, std::string* out, size_t out_len) {
  size_t ip = 0;
  out->clear();
  while (ip < in_len) {
    unsigned ctrl = in[ip++];
    if (ctrl < 32) {
      ctrl++;
      if (ip + ctrl > in_len) return false;
      out->append((const char*)in + ip, ctrl);
      ip += ctrl;
    } else {
      size_t len = ctrl >> 5;
      if (len == 7) {
        if (ip >= in_len) return false;
        len += in[ip++];
      }
      if (ip >= in_len) return false;
      const size_t ref_off = ((size_t)(ctrl & 0x1F) << 8) + in[ip++] + 1;
      if (ref_off > out->size()) return false;
      size_t ref = out->size() - ref_off;
      for (size_t i = 0; i < len + 2; ++i) out->push_back((*out)[ref + i]);
    }
  }
  return out->size() == out_len;
}

struct IO {
  std::string buf;
  size_t pos = 0;
  bool error = false;
  void op(int opcode) { rdb_put_len(buf, (uint64_t)opcode); }
  bool expect(int opcode) {
    uint64_t got = 0;
    if (!rdb_get_len(buf, pos, &got) || got != (uint64_t)opcode) {
      error = true;
      return false;
    }
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

static void api_SaveUnsigned(IO* io, uint64_t v) {
  io->op(kOpUint);
  rdb_put_len(io->buf, v);
}
static uint64_t api_LoadUnsigned(IO* io) {
  uint64_t v = 0;
  if (!io->expect(kOpUint) || !rdb_get_len(io->buf, io->pos, &v)) io->error = true;
  return v;
}
static void api_SaveDouble(IO* io, double v) {
  io->op(kOpDouble);
  io->buf.append((const char*)&v, 8);  // rdbSaveBinaryDoubleValue: IEEE 754, little-endian
}
static double api_LoadDouble(IO* io) {
  double v = 0;
  if (!io->expect(kOpDouble) || io->pos + 8 > io->buf.size()) {
    io->error = true;
    return 0;
  }
  std::memcpy(&v, io->buf.data() + io->pos, 8);
  io->pos += 8;
  return v;
}
static void api_SaveFloat(IO* io, float v) {
  io->op(kOpFloat);
  io->buf.append((const char*)&v, 4);  // rdbSaveBinaryFloatValue
}
static float api_LoadFloat(IO* io) {
  float v = 0;
  if (!io->expect(kOpFloat) || io->pos + 4 > io->buf.size()) {
    io->error = true;
    return 0;
  }
  std::memcpy(&v, io->buf.data() + io->pos, 4);
  io->pos += 4;
  return v;
}
static void api_SaveStringBuffer(IO* io, const char* p, size_t n) {
  io->op(kOpString);
  rdb_put_len(io->buf, n);
  io->buf.append(p, n);
}
static char* api_LoadStringBuffer(IO* io, size_t* lenptr) {
  std::string s;
  uint64_t len = 0;
  int enc = -1;
  bool ok = io->expect(kOpString);
  if (ok) {
    if (rdb_get_len(io->buf, io->pos, &len, &enc)) {
      ok = io->pos + len <= io->buf.size();
      if (ok) s.assign(io->buf, io->pos, len), io->pos += len;
    } else if (enc >= 0 && enc <= 2) {           // RDB_ENC_INT8 / INT16 / INT32: the string is the decimal of the integer
      const size_t n = (size_t)1 << enc;
      ok = io->pos + n <= io->buf.size();
      if (ok) {
        int64_t v = 0;
        if (n == 1) v = (int8_t)io->buf[io->pos];
        else if (n == 2) { int16_t t; std::memcpy(&t, io->buf.data() + io->pos, 2); v = t; }
        else { int32_t t; std::memcpy(&t, io->buf.data() + io->pos, 4); v = t; }
        io->pos += n;
        s = std::to_string(v);
      }
    } else if (enc == 3) {                       // RDB_ENC_LZF
      uint64_t clen = 0, ulen = 0;
      ok = rdb_get_len(io->buf, io->pos, &clen) && rdb_get_len(io->buf, io->pos, &ulen) && io->pos + clen <= io->buf.size() &&
           lzf_decompress((const unsigned char*)io->buf.data() + io->pos, clen, &s, ulen);
      if (ok) io->pos += clen;
    } else {
      ok = false;
    }
  }
  if (!ok) {
    io->error = true;
    if (lenptr) *lenptr = 0;
    return nullptr;
  }
  char* out = (char*)std::malloc(s.size() + 1);
  std::memcpy(out, s.data(), s.size());
  out[s.size()] = 0;
  if (lenptr) *lenptr = s.size();
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
    rdb_put_len(io.buf, module_type_id(kv.second.type->name, kv.second.type->encver));   // rdbSaveObject, RDB_TYPE_MODULE_2
    kv.second.type->m.rdb_save(&io, kv.second.value);
    io.op(kOpEof);
    put_s(out, kv.first);
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
      std::string key, payload;
      ok = take_s(in, pos, &key) && take_s(in, pos, &payload);
      if (!ok) break;
      IO io;
      io.buf = payload;
      uint64_t mid = 0;
      ok = rdb_get_len(io.buf, io.pos, &mid);   // the module type id names the data type and its encoding version
      RedisModuleType* mt = nullptr;
      for (auto& t : H.types)
        if (ok && (module_type_id(t.second->name, 0) >> 10) == (mid >> 10)) mt = t.second;
      if (!mt) {
        ok = false;
        break;
      }
      void* v = mt->m.rdb_load(&io, (int)(mid & 1023));
      if (!v || io.error || !io.expect(kOpEof) || io.pos != io.buf.size()) {
        ok = false;
        break;
      }
      auto it = H.types.find(mt->name);
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
