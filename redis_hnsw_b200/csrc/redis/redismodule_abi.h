/*
 * redismodule_abi.h — the slice of the Redis Modules C ABI this module binds.
 *
 * The reference reaches Redis through the `redis-module` crate (Cargo.toml:13), whose `redis_module!` macro
 * (src/lib.rs:498-514) expands to the exported `RedisModule_OnLoad` and whose `raw` module wraps the same function
 * table (src/types.rs:193-284 call raw::RedisModule_{Load,Save}{String,Unsigned,Double,Float} directly).  Neither the
 * crate nor redis' own `redismodule.h` is present in this image, so the handful of names, constants and signatures
 * needed are declared here from the public, stable module ABI (API version 1): a module receives a context whose
 * first word is the `GetApi(name, &fnptr)` resolver and looks every other entry point up by name.
 *
 * Only what hnsw_module.cpp calls is declared.  Every pointer is resolved in RedisModule_Init(); a name the host does
 * not provide fails the load instead of crashing later.
 */
#ifndef HNSW_REDISMODULE_ABI_H
#define HNSW_REDISMODULE_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REDISMODULE_OK 0
#define REDISMODULE_ERR 1
#define REDISMODULE_APIVER_1 1
#define REDISMODULE_READ (1 << 0)
#define REDISMODULE_WRITE (1 << 1)
#define REDISMODULE_KEYTYPE_EMPTY 0
#define REDISMODULE_KEYTYPE_MODULE 6
#define REDISMODULE_TYPE_METHOD_VERSION 1

typedef struct RedisModuleCtx RedisModuleCtx;
typedef struct RedisModuleKey RedisModuleKey;
typedef struct RedisModuleString RedisModuleString;
typedef struct RedisModuleIO RedisModuleIO;
typedef struct RedisModuleType RedisModuleType;
typedef struct RedisModuleDigest RedisModuleDigest;

typedef int (*RedisModuleCmdFunc)(RedisModuleCtx* ctx, RedisModuleString** argv, int argc);
typedef void* (*RedisModuleTypeLoadFunc)(RedisModuleIO* rdb, int encver);
typedef void (*RedisModuleTypeSaveFunc)(RedisModuleIO* rdb, void* value);
typedef void (*RedisModuleTypeRewriteFunc)(RedisModuleIO* aof, RedisModuleString* key, void* value);
typedef size_t (*RedisModuleTypeMemUsageFunc)(const void* value);
typedef void (*RedisModuleTypeDigestFunc)(RedisModuleDigest* digest, void* value);
typedef void (*RedisModuleTypeFreeFunc)(void* value);

/* Version-1 layout of the type-method table (the fields the reference fills at types.rs:157-174, 354-371). */
typedef struct RedisModuleTypeMethods {
  uint64_t version;
  RedisModuleTypeLoadFunc rdb_load;
  RedisModuleTypeSaveFunc rdb_save;
  RedisModuleTypeRewriteFunc aof_rewrite;
  RedisModuleTypeMemUsageFunc mem_usage;
  RedisModuleTypeDigestFunc digest;
  RedisModuleTypeFreeFunc free;
} RedisModuleTypeMethods;

/* Server events (Redis >= 6.0).  Optional: resolved if the host provides it (see RedisModule_Init below). */
typedef struct RedisModuleEvent {
  uint64_t id;
  uint64_t dataver;
} RedisModuleEvent;
typedef void (*RedisModuleEventCallback)(RedisModuleCtx* ctx, RedisModuleEvent eid, uint64_t subevent, void* data);
#define REDISMODULE_EVENT_PERSISTENCE 1
/* subevents 0..2 (3 on Redis >= 7.0) announce that a snapshot is about to START: RDB, AOF rewrite, sync RDB, sync AOF */
#define HNSW_PERSISTENCE_START_MAX_SUBEVENT 3

#define HNSW_RM_API(X)                                                                                              \
  X(void*, Alloc, (size_t bytes))                                                                                   \
  X(void, Free, (void* ptr))                                                                                        \
  X(int, CreateCommand,                                                                                             \
    (RedisModuleCtx * ctx, const char* name, RedisModuleCmdFunc cmdfunc, const char* strflags, int firstkey,       \
     int lastkey, int keystep))                                                                                     \
  X(void, SetModuleAttribs, (RedisModuleCtx * ctx, const char* name, int ver, int apiver))                          \
  X(int, IsModuleNameBusy, (const char* name))                                                                      \
  X(int, WrongArity, (RedisModuleCtx * ctx))                                                                        \
  X(void, AutoMemory, (RedisModuleCtx * ctx))                                                                       \
  X(RedisModuleType*, CreateDataType,                                                                               \
    (RedisModuleCtx * ctx, const char* name, int encver, RedisModuleTypeMethods* typemethods))                     \
  X(void*, OpenKey, (RedisModuleCtx * ctx, RedisModuleString * keyname, int mode))                                  \
  X(void, CloseKey, (RedisModuleKey * kp))                                                                          \
  X(int, KeyType, (RedisModuleKey * kp))                                                                            \
  X(int, DeleteKey, (RedisModuleKey * kp))                                                                          \
  X(RedisModuleType*, ModuleTypeGetType, (RedisModuleKey * kp))                                                     \
  X(void*, ModuleTypeGetValue, (RedisModuleKey * kp))                                                               \
  X(int, ModuleTypeSetValue, (RedisModuleKey * kp, RedisModuleType * mt, void* value))                              \
  X(RedisModuleString*, CreateString, (RedisModuleCtx * ctx, const char* ptr, size_t len))                          \
  X(void, FreeString, (RedisModuleCtx * ctx, RedisModuleString * str))                                              \
  X(const char*, StringPtrLen, (const RedisModuleString* str, size_t* len))                                         \
  X(int, ReplyWithError, (RedisModuleCtx * ctx, const char* err))                                                   \
  X(int, ReplyWithSimpleString, (RedisModuleCtx * ctx, const char* msg))                                            \
  X(int, ReplyWithLongLong, (RedisModuleCtx * ctx, long long ll))                                                   \
  X(int, ReplyWithDouble, (RedisModuleCtx * ctx, double d))                                                         \
  X(int, ReplyWithArray, (RedisModuleCtx * ctx, long len))                                                          \
  X(int, ReplyWithStringBuffer, (RedisModuleCtx * ctx, const char* buf, size_t len))                                \
  X(int, ReplyWithNull, (RedisModuleCtx * ctx))                                                                     \
  X(void, SaveUnsigned, (RedisModuleIO * io, uint64_t value))                                                       \
  X(uint64_t, LoadUnsigned, (RedisModuleIO * io))                                                                   \
  X(void, SaveDouble, (RedisModuleIO * io, double value))                                                           \
  X(double, LoadDouble, (RedisModuleIO * io))                                                                       \
  X(void, SaveFloat, (RedisModuleIO * io, float value))                                                             \
  X(float, LoadFloat, (RedisModuleIO * io))                                                                         \
  X(void, SaveStringBuffer, (RedisModuleIO * io, const char* str, size_t len))                                      \
  X(char*, LoadStringBuffer, (RedisModuleIO * io, size_t* lenptr))

#ifdef HNSW_REDISMODULE_MAIN
#define HNSW_RM_DECL(ret, name, args) ret(*RedisModule_##name) args = 0;
#else
#define HNSW_RM_DECL(ret, name, args) extern ret(*RedisModule_##name) args;
#endif
HNSW_RM_API(HNSW_RM_DECL)
#undef HNSW_RM_DECL
#ifdef HNSW_REDISMODULE_MAIN
int (*RedisModule_SubscribeToServerEvent)(RedisModuleCtx* ctx, RedisModuleEvent event, RedisModuleEventCallback callback) = 0;
#else
extern int (*RedisModule_SubscribeToServerEvent)(RedisModuleCtx* ctx, RedisModuleEvent event, RedisModuleEventCallback callback);
#endif

#ifdef HNSW_REDISMODULE_MAIN
/* What redismodule.h's RedisModule_Init does: the first word of the context is the GetApi resolver. */
static int RedisModule_Init(RedisModuleCtx* ctx, const char* name, int ver, int apiver) {
  typedef int (*GetApiFn)(const char*, void*);
  GetApiFn get_api = (GetApiFn)((void**)ctx)[0];
  if (!get_api) return REDISMODULE_ERR;
#define HNSW_RM_GET(ret, fname, args) \
  if (get_api("RedisModule_" #fname, (void*)&RedisModule_##fname) != REDISMODULE_OK || !RedisModule_##fname) return REDISMODULE_ERR;
  HNSW_RM_API(HNSW_RM_GET)
#undef HNSW_RM_GET
  if (get_api("RedisModule_SubscribeToServerEvent", (void*)&RedisModule_SubscribeToServerEvent) != REDISMODULE_OK)
    RedisModule_SubscribeToServerEvent = 0; /* optional */
  if (RedisModule_IsModuleNameBusy(name)) return REDISMODULE_ERR;
  RedisModule_SetModuleAttribs(ctx, name, ver, apiver);
  return REDISMODULE_OK;
}
#endif

#ifdef __cplusplus
}
#endif
#endif /* HNSW_REDISMODULE_ABI_H */
