/* napi_min.h -- hand-declared subset of Node's stable N-API (node_api.h / js_native_api.h, ABI
 * version 6) used by rt2015_napi.c, so the addon can be syntax- and type-checked in an image that
 * has no Node.js headers.  Signatures follow the public N-API documentation; never linked. */
#ifndef NAPI_MIN_H
#define NAPI_MIN_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_generic_failure = 9 } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external,
               napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
               napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
#define NAPI_AUTO_LENGTH ((size_t)-1)

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_value_int64(napi_env env, napi_value value, int64_t* result);
napi_status napi_get_value_bigint_uint64(napi_env env, napi_value value, uint64_t* result, bool* lossless);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char* buf, size_t bufsize, size_t* result);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_create_uint32(napi_env env, uint32_t value, napi_value* result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_create_double(napi_env env, double value, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_get_named_property(napi_env env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_has_named_property(napi_env env, napi_value object, const char* utf8name, bool* result);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result);
napi_status napi_create_bigint_uint64(napi_env env, uint64_t value, napi_value* result);
napi_status napi_is_typedarray(napi_env env, napi_value value, bool* result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_arraybuffer_info(napi_env env, napi_value arraybuffer, void** data, size_t* byte_length);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
#endif
