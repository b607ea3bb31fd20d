#define HAVE_XXHASH_H 0
