/* minimal stand-in for <jni.h>: only what flinkms_jni.c uses, to syntax- and type-check the shim without a JDK */
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef double jdouble; typedef unsigned char jboolean; typedef signed char jbyte; typedef jint jsize;
typedef void *jobject; typedef jobject jclass; typedef jobject jlongArray; typedef jobject jdoubleArray; typedef jobject jintArray; typedef jobject jarray;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv *, const char *);
    jsize (*GetArrayLength)(JNIEnv *, jarray);
    void *(*GetDirectBufferAddress)(JNIEnv *, jobject);
    void (*GetIntArrayRegion)(JNIEnv *, jintArray, jsize, jsize, jint *);
    jdoubleArray (*NewDoubleArray)(JNIEnv *, jsize);
    jlongArray (*NewLongArray)(JNIEnv *, jsize);
    void (*SetDoubleArrayRegion)(JNIEnv *, jdoubleArray, jsize, jsize, const jdouble *);
    void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
    jint (*ThrowNew)(JNIEnv *, jclass, const char *);
};
#define JNIEXPORT
#define JNICALL
