/* napi_mock.c -- TEST INFRASTRUCTURE.  A minimal in-process implementation of the N-API subset declared in
 * ../napi_min.h (numbers, BigInts, strings, ArrayBuffers / TypedArrays, plain objects, functions, pending
 * exceptions), so that the addon's marshalling code (../rt2015_napi.c) can be EXECUTED in an image that has no
 * Node.js: addon_test.c registers the addon through rt2015_init() and calls its exports the way JavaScript would.
 * Values are never freed (short-lived test process). */
#include "napi_mock.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static napi_value new_value(mock_kind k) {
    napi_value v = (napi_value)calloc(1, sizeof *v);
    v->kind = k;
    return v;
}
napi_value mock_undefined(void) { return new_value(MOCK_UNDEFINED); }
napi_value mock_null(void) { return new_value(MOCK_NULL); }
napi_value mock_number(double d) { napi_value v = new_value(MOCK_NUMBER); v->num = d; return v; }
napi_value mock_bigint(uint64_t u) { napi_value v = new_value(MOCK_BIGINT); v->big = u; return v; }
napi_value mock_string(const char* s) { napi_value v = new_value(MOCK_STRING); v->str = strdup(s); return v; }
napi_value mock_object(void) { return new_value(MOCK_OBJECT); }
static size_t elem_size(napi_typedarray_type t) {
    static const size_t esz[] = {1, 1, 1, 2, 2, 4, 4, 4, 8, 8, 8};
    return esz[t];
}
napi_value mock_typedarray(napi_typedarray_type t, size_t length, const void* init) {
    napi_value ab = new_value(MOCK_ARRAYBUFFER);
    ab->byte_length = length * elem_size(t);
    ab->data = calloc(ab->byte_length ? ab->byte_length : 1, 1);
    if (init && ab->byte_length) memcpy(ab->data, init, ab->byte_length);
    napi_value v = new_value(MOCK_TYPEDARRAY);
    v->ttype = t; v->length = length; v->data = ab->data; v->buffer = ab;
    return v;
}
napi_value mock_get(napi_value obj, const char* key) {
    for (mock_prop* p = obj->props; p; p = p->next)
        if (!strcmp(p->key, key)) return p->value;
    return NULL;
}
void mock_set(napi_value obj, const char* key, napi_value val) {
    for (mock_prop* p = obj->props; p; p = p->next)
        if (!strcmp(p->key, key)) { p->value = val; return; }
    mock_prop* p = (mock_prop*)calloc(1, sizeof *p);
    p->key = strdup(key); p->value = val; p->next = obj->props; obj->props = p;
}
napi_value mock_call(napi_env env, napi_value exports, const char* name, size_t argc, napi_value* argv) {
    napi_value fn = mock_get(exports, name);
    env->pending = 0;
    env->message[0] = 0;
    if (!fn || fn->kind != MOCK_FUNCTION) {
        env->pending = 1;
        snprintf(env->message, sizeof env->message, "TypeError: %s is not a function", name);
        return NULL;
    }
    struct napi_callback_info__ info = {argc, argv};
    napi_value r = fn->fn(env, &info);
    return env->pending ? NULL : (r ? r : mock_undefined());
}

/* ---- the N-API subset ------------------------------------------------------------------------- */
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
    (void)env; (void)this_arg; (void)data;
    size_t want = *argc;
    for (size_t i = 0; i < want; i++) argv[i] = i < cbinfo->argc ? cbinfo->argv[i] : mock_undefined();
    *argc = cbinfo->argc;
    return napi_ok;
}
static napi_status throw_(napi_env env, const char* prefix, const char* msg) {
    env->pending = 1;
    snprintf(env->message, sizeof env->message, "%s%s", prefix, msg ? msg : "");
    return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char* code, const char* msg) { (void)code; return throw_(env, "Error: ", msg); }
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg) { (void)code; return throw_(env, "TypeError: ", msg); }
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result) {
    (void)env;
    switch (value->kind) {
        case MOCK_UNDEFINED: *result = napi_undefined; break;
        case MOCK_NULL: *result = napi_null; break;
        case MOCK_NUMBER: *result = napi_number; break;
        case MOCK_BIGINT: *result = napi_bigint; break;
        case MOCK_STRING: *result = napi_string; break;
        case MOCK_FUNCTION: *result = napi_function; break;
        default: *result = napi_object; break;
    }
    return napi_ok;
}
#define NEED(v, k) do { if (!(v) || (v)->kind != (k)) return napi_invalid_arg; } while (0)
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result) { (void)env; NEED(value, MOCK_NUMBER); *result = (int32_t)value->num; return napi_ok; }
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result) { (void)env; NEED(value, MOCK_NUMBER); *result = (uint32_t)value->num; return napi_ok; }
napi_status napi_get_value_int64(napi_env env, napi_value value, int64_t* result) { (void)env; NEED(value, MOCK_NUMBER); *result = (int64_t)value->num; return napi_ok; }
napi_status napi_get_value_double(napi_env env, napi_value value, double* result) { (void)env; NEED(value, MOCK_NUMBER); *result = value->num; return napi_ok; }
napi_status napi_get_value_bigint_uint64(napi_env env, napi_value value, uint64_t* result, bool* lossless) {
    (void)env; NEED(value, MOCK_BIGINT); *result = value->big; if (lossless) *lossless = true; return napi_ok;
}
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char* buf, size_t bufsize, size_t* result) {
    (void)env; NEED(value, MOCK_STRING);
    size_t n = strlen(value->str);
    if (buf && bufsize) { if (n > bufsize - 1) n = bufsize - 1; memcpy(buf, value->str, n); buf[n] = 0; }
    if (result) *result = n;
    return napi_ok;
}
napi_status napi_create_uint32(napi_env env, uint32_t value, napi_value* result) { (void)env; *result = mock_number(value); return napi_ok; }
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result) { (void)env; *result = mock_number(value); return napi_ok; }
napi_status napi_create_double(napi_env env, double value, napi_value* result) { (void)env; *result = mock_number(value); return napi_ok; }
napi_status napi_create_bigint_uint64(napi_env env, uint64_t value, napi_value* result) { (void)env; *result = mock_bigint(value); return napi_ok; }
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result) {
    (void)env;
    if (length == NAPI_AUTO_LENGTH) { *result = mock_string(str); return napi_ok; }
    char* tmp = (char*)calloc(length + 1, 1);
    memcpy(tmp, str, length);
    *result = mock_string(tmp);
    free(tmp);
    return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value* result) { (void)env; *result = mock_object(); return napi_ok; }
napi_status napi_get_named_property(napi_env env, napi_value object, const char* utf8name, napi_value* result) {
    (void)env;
    if (!object || (object->kind != MOCK_OBJECT && object->kind != MOCK_FUNCTION)) return napi_object_expected;
    napi_value v = mock_get(object, utf8name);
    *result = v ? v : mock_undefined();
    return napi_ok;
}
napi_status napi_has_named_property(napi_env env, napi_value object, const char* utf8name, bool* result) {
    (void)env;
    if (!object || object->kind != MOCK_OBJECT) return napi_object_expected;
    *result = mock_get(object, utf8name) != NULL;
    return napi_ok;
}
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value) {
    (void)env;
    if (!object || object->kind != MOCK_OBJECT) return napi_object_expected;
    mock_set(object, utf8name, value);
    return napi_ok;
}
napi_status napi_is_typedarray(napi_env env, napi_value value, bool* result) { (void)env; *result = value && value->kind == MOCK_TYPEDARRAY; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset) {
    (void)env; NEED(typedarray, MOCK_TYPEDARRAY);
    if (type) *type = typedarray->ttype;
    if (length) *length = typedarray->length;
    if (data) *data = typedarray->data;
    if (arraybuffer) *arraybuffer = typedarray->buffer;
    if (byte_offset) *byte_offset = 0;
    return napi_ok;
}
napi_status napi_get_arraybuffer_info(napi_env env, napi_value arraybuffer, void** data, size_t* byte_length) {
    (void)env; NEED(arraybuffer, MOCK_ARRAYBUFFER);
    if (data) *data = arraybuffer->data;
    if (byte_length) *byte_length = arraybuffer->byte_length;
    return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result) {
    (void)env;
    napi_value ab = new_value(MOCK_ARRAYBUFFER);
    ab->byte_length = byte_length;
    ab->data = calloc(byte_length ? byte_length : 1, 1);
    if (data) *data = ab->data;
    *result = ab;
    return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result) {
    (void)env; NEED(arraybuffer, MOCK_ARRAYBUFFER);
    if (byte_offset + length * elem_size(type) > arraybuffer->byte_length) return napi_invalid_arg;
    napi_value v = new_value(MOCK_TYPEDARRAY);
    v->ttype = type; v->length = length; v->data = (char*)arraybuffer->data + byte_offset; v->buffer = arraybuffer;
    *result = v;
    return napi_ok;
}
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result) {
    (void)env; (void)utf8name; (void)length; (void)data;
    napi_value v = new_value(MOCK_FUNCTION);
    v->fn = cb;
    *result = v;
    return napi_ok;
}
