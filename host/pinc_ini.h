/* pinc_ini.h — minimal .ini reader for the C host of libpinc_b200 with the semantics PINC relies on:
 * iniparser 3.1 dictionary ("section:key", keys lower-cased; /root/reference/lib/iniparser/src/iniparser.c:376),
 * ';' and '#' comments, comma lists cyclically expanded to the requested length and parsed with atof
 * (/root/reference/src/io.c:332-433, 823), "section:key=value" command-line overrides (io.c:273-276), and the
 * "pc"/"tot" suffixes of units.c:138-158.  Written from scratch; iniparser itself is not linked. */
#ifndef PINC_INI_H
#define PINC_INI_H

typedef struct { char *key, *val; } IniEntry;
typedef struct { IniEntry *e; int n, cap; } Ini;

Ini *iniLoad(const char *path);                         /* NULL if the file cannot be read */
void iniFree(Ini *ini);
void iniSet(Ini *ini, const char *key, const char *val);           /* key "section:name", any case */
int  iniApplyOverride(Ini *ini, const char *arg);                  /* "section:key=value"; 0 if malformed */
const char *iniRaw(const Ini *ini, const char *key);               /* aborts like msg(ERROR) if missing */
int  iniHas(const Ini *ini, const char *key);
int  iniNElements(const Ini *ini, const char *key);
void iniGetStr(const Ini *ini, const char *key, int i, int n, char *out, int outlen);  /* element i of n, cyclic */
void iniGetDoubles(const Ini *ini, const char *key, int n, double *out);
void iniGetInts(const Ini *ini, const char *key, int n, int *out);
void iniGetLongs(const Ini *ini, const char *key, int n, long *out);
double iniGetDouble(const Ini *ini, const char *key);
int  iniGetInt(const Ini *ini, const char *key);
void iniSetDoubles(Ini *ini, const char *key, int n, const double *v);   /* hex floats, like iniSetDoubleArr */
void iniApplySuffix(Ini *ini, const char *key, const char *suffix, const double *mul, int nMul);

#endif
