/*
 * oracle/ora_ini.c -- restatement of the input.inf reader (TEST INFRASTRUCTURE, see ora.h).
 *
 * Follows src/shared/m_readini.f90:29-100 (readini_c) and :103-170 (typed wrappers) and
 * src/shared/m_system.f90:75-102 (${ENV} expansion).
 *
 * Quirks kept on purpose:
 *   - per key, the file is scanned from the top; the first non-comment line whose left-trimmed
 *     text STARTS WITH the key (prefix match, m_readini.f90:83) and whose next non-blank
 *     character is '=' (:87) wins;
 *   - lines starting with '#' or '!' are comments (:80);
 *   - the value is parsed by a Fortran list-directed read (:89): a quoted string, or an
 *     undelimited token that ends at blank, comma, slash or end of line;
 *   - missing keys fall back to the caller's default unless strict_mode (:64-75).
 */
#include "ora.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static ora_ini *ini_alloc(void) {
    ora_ini *ini = (ora_ini *)calloc(1, sizeof(ora_ini));
    return ini;
}

static void ini_push(ora_ini *ini, const char *s, size_t n) {
    ini->lines = (char **)realloc(ini->lines, sizeof(char *) * (size_t)(ini->nlines + 1));
    char *l = (char *)malloc(n + 1);
    memcpy(l, s, n);
    l[n] = 0;
    /* strip CR */
    if (n > 0 && l[n - 1] == '\r') l[n - 1] = 0;
    ini->lines[ini->nlines++] = l;
}

ora_ini *ora_ini_from_text(const char *text) {
    ora_ini *ini = ini_alloc();
    const char *p = text;
    while (*p) {
        const char *e = strchr(p, '\n');
        size_t n = e ? (size_t)(e - p) : strlen(p);
        ini_push(ini, p, n);
        if (!e) break;
        p = e + 1;
    }
    return ini;
}

ora_ini *ora_ini_open(const char *path) {
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    size_t nr = fread(buf, 1, (size_t)sz, fp);
    buf[nr] = 0;
    fclose(fp);
    ora_ini *ini = ora_ini_from_text(buf);
    free(buf);
    return ini;
}

void ora_ini_close(ora_ini *ini) {
    if (!ini) return;
    for (int i = 0; i < ini->nlines; i++) free(ini->lines[i]);
    free(ini->lines);
    free(ini);
}

/* Fortran list-directed read of ONE character item from s (m_readini.f90:89 "read(cline,*) var") */
static void list_directed_char(const char *s, char *out, size_t cap) {
    size_t n = 0;
    while (*s == ' ' || *s == '\t') s++;
    if (*s == '\'' || *s == '"') {
        char q = *s++;
        while (*s) {
            if (*s == q) {
                if (s[1] == q) { /* doubled delimiter = literal */
                    if (n + 1 < cap) out[n++] = q;
                    s += 2;
                    continue;
                }
                break;
            }
            if (n + 1 < cap) out[n++] = *s;
            s++;
        }
    } else {
        while (*s && *s != ' ' && *s != '\t' && *s != ',' && *s != '/') {
            if (n + 1 < cap) out[n++] = *s;
            s++;
        }
    }
    out[n] = 0;
}

/* m_system.f90:75-102 -- expand ${VAR}; restated for well-formed input */
static void expenv(char *str, size_t cap) {
    char res[ORA_STRLEN * 2];
    size_t n = 0;
    const char *p = str;
    while (*p) {
        if (p[0] == '$' && p[1] == '{') {
            const char *e = strchr(p, '}');
            if (!e) break;
            char name[ORA_STRLEN];
            size_t ln = (size_t)(e - (p + 2));
            if (ln >= sizeof(name)) ln = sizeof(name) - 1;
            memcpy(name, p + 2, ln);
            name[ln] = 0;
            const char *v = getenv(name);
            if (v) {
                size_t lv = strlen(v);
                if (n + lv < sizeof(res)) { memcpy(res + n, v, lv); n += lv; }
            }
            p = e + 1;
        } else {
            if (n + 1 < sizeof(res)) res[n++] = *p;
            p++;
        }
    }
    while (*p) { if (n + 1 < sizeof(res)) res[n++] = *p; p++; }
    res[n] = 0;
    /* trim trailing blanks as trim() does */
    while (n > 0 && res[n - 1] == ' ') res[--n] = 0;
    strncpy(str, res, cap - 1);
    str[cap - 1] = 0;
}

int ora_readini_c(const ora_ini *ini, const char *key, char *var, const char *def) {
    char keyword[ORA_STRLEN];
    /* keyword = trim(adjustl(key)) */
    while (*key == ' ') key++;
    strncpy(keyword, key, sizeof(keyword) - 1);
    keyword[sizeof(keyword) - 1] = 0;
    size_t keylen = strlen(keyword);
    while (keylen > 0 && keyword[keylen - 1] == ' ') keyword[--keylen] = 0;

    for (int l = 0; ini && l < ini->nlines; l++) {
        const char *c = ini->lines[l];
        while (*c == ' ' || *c == '\t') c++; /* adjustl */
        if (*c == '#' || *c == '!') continue;
        if (strncmp(c, keyword, keylen) == 0) {
            const char *q = c + keylen;
            while (*q == ' ' || *q == '\t') q++;
            if (*q == '=') {
                q++;
                list_directed_char(q, var, ORA_STRLEN);
                expenv(var, ORA_STRLEN);
                return 1;
            }
        }
    }
    if (ini && ini->strict_mode) {
        fprintf(stderr, "[ora readini] key %s is not found. Program terminate ...\n", keyword);
        exit(1);
    }
    if (ini && ini->verbose)
        fprintf(stderr, "[ora readini] key %s is not found. Use default value %s instead.\n", keyword, def);
    strncpy(var, def, ORA_STRLEN - 1);
    var[ORA_STRLEN - 1] = 0;
    expenv(var, ORA_STRLEN);
    return 0;
}

/* Fortran real literal -> C: allow d/D/q exponents and a trailing list separator */
static void fortran_real_token(const char *s, char *tok, size_t cap) {
    size_t n = 0;
    while (*s == ' ' || *s == '\t') s++;
    while (*s && *s != ' ' && *s != '\t' && *s != ',' && *s != '/' && n + 1 < cap) {
        char ch = *s++;
        if (ch == 'd' || ch == 'D' || ch == 'q' || ch == 'Q') ch = 'e';
        tok[n++] = ch;
    }
    tok[n] = 0;
}

int ora_readini_d(const ora_ini *ini, const char *key, double *var, double def) {
    char avar[ORA_STRLEN], adef[ORA_STRLEN], tok[ORA_STRLEN];
    snprintf(adef, sizeof(adef), "%.17g", def);
    int found = ora_readini_c(ini, key, avar, adef);
    fortran_real_token(avar, tok, sizeof(tok));
    *var = strtod(tok, NULL);
    return found;
}

int ora_readini_s(const ora_ini *ini, const char *key, float *var, float def) {
    char avar[ORA_STRLEN], adef[ORA_STRLEN], tok[ORA_STRLEN];
    snprintf(adef, sizeof(adef), "%.9g", (double)def);
    int found = ora_readini_c(ini, key, avar, adef);
    fortran_real_token(avar, tok, sizeof(tok));
    *var = strtof(tok, NULL);
    return found;
}

int ora_readini_i(const ora_ini *ini, const char *key, int *var, int def) {
    char avar[ORA_STRLEN], adef[ORA_STRLEN];
    snprintf(adef, sizeof(adef), "%d", def);
    int found = ora_readini_c(ini, key, avar, adef);
    *var = (int)strtol(avar, NULL, 10);
    return found;
}

int ora_readini_l(const ora_ini *ini, const char *key, int *var, int def) {
    char avar[ORA_STRLEN];
    int found = ora_readini_c(ini, key, avar, def ? "T" : "F");
    const char *p = avar;
    while (*p == ' ') p++;
    if (*p == '.') p++;
    *var = (*p == 'T' || *p == 't') ? 1 : 0;
    return found;
}
