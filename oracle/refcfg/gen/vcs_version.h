#define DAV1D_VERSION "1.0.0-oracle"
