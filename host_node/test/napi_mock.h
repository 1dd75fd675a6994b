/* napi_mock.h -- TEST INFRASTRUCTURE: value model of the in-process N-API stand-in (see napi_mock.c). */
#ifndef NAPI_MOCK_H
#define NAPI_MOCK_H
#include "../napi_min.h"

typedef enum { MOCK_UNDEFINED, MOCK_NULL, MOCK_NUMBER, MOCK_BIGINT, MOCK_STRING, MOCK_ARRAYBUFFER, MOCK_TYPEDARRAY, MOCK_OBJECT,
               MOCK_FUNCTION } mock_kind;
typedef struct mock_prop { char* key; napi_value value; struct mock_prop* next; } mock_prop;
struct napi_value__ {
    mock_kind kind;
    double num;
    uint64_t big;
    char* str;
    napi_typedarray_type ttype;
    size_t length, byte_length;
    void* data;
    napi_value buffer;
    mock_prop* props;
    napi_callback fn;
};
struct napi_env__ { int pending; char message[700]; };
struct napi_callback_info__ { size_t argc; napi_value* argv; };

napi_value mock_undefined(void);
napi_value mock_null(void);
napi_value mock_number(double d);
napi_value mock_bigint(uint64_t u);
napi_value mock_string(const char* s);
napi_value mock_object(void);
napi_value mock_typedarray(napi_typedarray_type t, size_t length, const void* init);
napi_value mock_get(napi_value obj, const char* key);
void mock_set(napi_value obj, const char* key, napi_value val);
/* exports.name(argv...) -- NULL with env->pending set when the call threw */
napi_value mock_call(napi_env env, napi_value exports, const char* name, size_t argc, napi_value* argv);
#endif
