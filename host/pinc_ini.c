/* pinc_ini.c — see pinc_ini.h */
#define _POSIX_C_SOURCE 200809L
#include "pinc_ini.h"
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void die(const char *fmt, const char *a){
	fprintf(stderr, "ERROR: ");
	fprintf(stderr, fmt, a);
	fprintf(stderr, "\n");
	exit(EXIT_FAILURE);
}
static char *dupLower(const char *s){
	char *r = malloc(strlen(s)+1);
	size_t i;
	for(i = 0; s[i]; i++) r[i] = (char)tolower((unsigned char)s[i]);
	r[i] = 0;
	return r;
}
static char *trim(char *s){
	while(*s && isspace((unsigned char)*s)) s++;
	char *e = s + strlen(s);
	while(e > s && isspace((unsigned char)e[-1])) *--e = 0;
	return s;
}
void iniSet(Ini *ini, const char *key, const char *val){
	char *k = dupLower(key);
	for(int i = 0; i < ini->n; i++) if(!strcmp(ini->e[i].key, k)){
		free(ini->e[i].val); ini->e[i].val = strdup(val); free(k); return;
	}
	if(ini->n == ini->cap){ ini->cap = ini->cap ? 2*ini->cap : 64; ini->e = realloc(ini->e, ini->cap*sizeof(IniEntry)); }
	ini->e[ini->n].key = k; ini->e[ini->n].val = strdup(val); ini->n++;
}
Ini *iniLoad(const char *path){
	FILE *f = fopen(path, "r");
	if(!f) return NULL;
	Ini *ini = calloc(1, sizeof *ini);
	char line[4096], sec[256] = "", key[512];
	while(fgets(line, sizeof line, f)){
		char *s = trim(line);
		if(!*s || *s == ';' || *s == '#') continue;
		if(*s == '['){
			char *e = strchr(s, ']');
			if(e){ *e = 0; snprintf(sec, sizeof sec, "%s", trim(s+1)); }
			continue;
		}
		char *eq = strchr(s, '=');
		if(!eq) continue;
		*eq = 0;
		char *val = eq + 1;
		for(char *c = val; *c; c++) if(*c == ';' || *c == '#'){ *c = 0; break; }
		snprintf(key, sizeof key, "%s:%s", sec, trim(s));
		iniSet(ini, key, trim(val));
	}
	fclose(f);
	return ini;
}
void iniFree(Ini *ini){
	if(!ini) return;
	for(int i = 0; i < ini->n; i++){ free(ini->e[i].key); free(ini->e[i].val); }
	free(ini->e); free(ini);
}
int iniApplyOverride(Ini *ini, const char *arg){
	const char *eq = strchr(arg, '=');
	if(!eq || !strchr(arg, ':') || strchr(arg, ':') > eq) return 0;
	char key[512];
	snprintf(key, sizeof key, "%.*s", (int)(eq-arg), arg);
	iniSet(ini, key, eq+1);
	return 1;
}
int iniHas(const Ini *ini, const char *key){
	char *k = dupLower(key);
	int found = 0;
	for(int i = 0; i < ini->n && !found; i++) found = !strcmp(ini->e[i].key, k);
	free(k);
	return found;
}
const char *iniRaw(const Ini *ini, const char *key){
	char *k = dupLower(key);
	for(int i = 0; i < ini->n; i++) if(!strcmp(ini->e[i].key, k)){ free(k); return ini->e[i].val; }
	die("Key \"%s\" not found in input file", key);       /* src/io.c:316-321 */
	return NULL;
}
int iniNElements(const Ini *ini, const char *key){
	const char *s = iniRaw(ini, key);
	if(!*s) return 0;
	int n = 1;
	for(; *s; s++) if(*s == ',') n++;
	return n;
}
void iniGetStr(const Ini *ini, const char *key, int i, int n, char *out, int outlen){
	(void)n;
	const char *s = iniRaw(ini, key);
	int have = iniNElements(ini, key);
	if(have < 1){ out[0] = 0; return; }
	int want = i % have, cur = 0;
	const char *b = s;
	for(const char *p = s; ; p++){
		if(*p == ',' || !*p){
			if(cur == want){
				char tmp[512];
				snprintf(tmp, sizeof tmp, "%.*s", (int)(p-b), b);
				snprintf(out, outlen, "%s", trim(tmp));
				return;
			}
			cur++; b = p+1;
			if(!*p) break;
		}
	}
	out[0] = 0;
}
/* atof semantics: numeric prefix (decimal or hex float), else 0 */
static double atofPrefix(const char *s){ return strtod(s, NULL); }
void iniGetDoubles(const Ini *ini, const char *key, int n, double *out){
	char tok[512];
	for(int i = 0; i < n; i++){ iniGetStr(ini, key, i, n, tok, sizeof tok); out[i] = atofPrefix(tok); }
}
void iniGetInts(const Ini *ini, const char *key, int n, int *out){
	char tok[512];
	for(int i = 0; i < n; i++){ iniGetStr(ini, key, i, n, tok, sizeof tok); out[i] = (int)atofPrefix(tok); }
}
void iniGetLongs(const Ini *ini, const char *key, int n, long *out){
	char tok[512];
	for(int i = 0; i < n; i++){ iniGetStr(ini, key, i, n, tok, sizeof tok); out[i] = (long)atofPrefix(tok); }
}
double iniGetDouble(const Ini *ini, const char *key){ return atofPrefix(iniRaw(ini, key)); }
int iniGetInt(const Ini *ini, const char *key){ return (int)atofPrefix(iniRaw(ini, key)); }
void iniSetDoubles(Ini *ini, const char *key, int n, const double *v){
	char buf[4096]; int off = 0;
	for(int i = 0; i < n; i++) off += snprintf(buf+off, sizeof buf - off, "%s%a", i ? "," : "", v[i]);
	iniSet(ini, key, buf);
}
void iniApplySuffix(Ini *ini, const char *key, const char *suffix, const double *mul, int nMul){
	if(!iniHas(ini, key)) return;
	int n = iniNElements(ini, key);
	double *v = malloc(sizeof(double)*(n > 0 ? n : 1));
	char tok[512];
	for(int i = 0; i < n; i++){
		iniGetStr(ini, key, i, n, tok, sizeof tok);
		v[i] = atofPrefix(tok);
		if(strstr(tok, suffix)) v[i] *= mul[i % nMul];
	}
	iniSetDoubles(ini, key, n, v);
	free(v);
}
