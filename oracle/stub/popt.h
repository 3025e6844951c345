/*
 * oracle/stub/popt.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for libpopt so that the UNMODIFIED reference translation
 * unit (/root/reference/main-cli.c, which does `#include <popt.h>` at line 22)
 * compiles in a container that has neither the popt header nor libpopt.so.
 * Only the symbols main-cli.c:1243-1401 touches are provided.  popt parses
 * argv and nothing else: no arithmetic of the SpMV path lives in it, so this
 * stub cannot change any number the reference produces.
 *
 * Written from the popt(3) man-page semantics, not from popt sources.
 */
#ifndef SMVP_ORACLE_STUB_POPT_H
#define SMVP_ORACLE_STUB_POPT_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <errno.h>

#define POPT_ARG_NONE 0
#define POPT_ARG_STRING 1
#define POPT_ARG_INT 2
#define POPT_ARG_INCLUDE_TABLE 4

#define POPT_CONTEXT_NO_EXEC (1 << 0)
#define POPT_CONTEXT_POSIXMEHARDER (1 << 2)

#define POPT_ERROR_NOARG -10
#define POPT_ERROR_BADOPT -11
#define POPT_ERROR_BADQUOTE -15
#define POPT_ERROR_BADNUMBER -17
#define POPT_ERROR_OVERFLOW -18
#define POPT_ERROR_NULLARG -20

#define POPT_BADOPTION_NOALIAS (1 << 0)

struct poptOption
{
    const char *longName;
    char shortName;
    int argInfo;
    void *arg;
    int val;
    const char *descrip;
    const char *argDescrip;
};

#define POPT_AUTOHELP {NULL, '\0', POPT_ARG_INCLUDE_TABLE, NULL, 0, "Help options:", NULL},
#define POPT_TABLEEND {NULL, '\0', 0, NULL, 0, NULL, NULL}

typedef struct smvp_stub_popt_ctx
{
    int argc;
    const char **argv;
    const struct poptOption *table;
    int next;        /* next argv slot to look at            */
    int optsDone;    /* saw first positional (POSIXMEHARDER) */
    const char *bad; /* option that caused the last error    */
    const char *help;
    const char *cluster; /* rest of a "-abc" short cluster   */
} *poptContext;

static inline poptContext poptGetContext(const char *name, int argc, const char **argv,
                                         const struct poptOption *options, unsigned int flags)
{
    (void)name;
    (void)flags;
    poptContext c = (poptContext)calloc(1, sizeof(*c));
    c->argc = argc;
    c->argv = argv;
    c->table = options;
    c->next = 1;
    return c;
}

static inline void poptSetOtherOptionHelp(poptContext c, const char *text) { c->help = text; }

static inline void poptPrintUsage(poptContext c, FILE *fp, int flags)
{
    (void)flags;
    fprintf(fp, "Usage: %s %s\n", c->argv[0], c->help ? c->help : "[OPTION...]");
}

static inline const struct poptOption *smvp_stub_popt_find(poptContext c, const char *lname, size_t llen, char sname)
{
    const struct poptOption *o;
    for (o = c->table; o->longName || o->shortName || o->argInfo; o++)
    {
        if (o->argInfo == POPT_ARG_INCLUDE_TABLE)
            continue;
        if (lname && o->longName && strlen(o->longName) == llen && strncmp(o->longName, lname, llen) == 0)
            return o;
        if (!lname && sname && o->shortName == sname)
            return o;
    }
    return NULL;
}

static inline int smvp_stub_popt_store(poptContext c, const struct poptOption *o, const char *optarg_, const char *self)
{
    if (o->argInfo == POPT_ARG_NONE)
        return o->val;
    if (optarg_ == NULL)
    {
        if (c->next >= c->argc)
        {
            c->bad = self;
            return POPT_ERROR_NOARG;
        }
        optarg_ = c->argv[c->next++];
    }
    if (o->argInfo == POPT_ARG_STRING)
    {
        if (o->arg)
            *(const char **)o->arg = optarg_;
    }
    else if (o->argInfo == POPT_ARG_INT)
    {
        char *end = NULL;
        errno = 0;
        long v = strtol(optarg_, &end, 0);
        if (end == optarg_ || *end != '\0')
        {
            c->bad = self;
            return POPT_ERROR_BADNUMBER;
        }
        if (errno == ERANGE || v > INT_MAX || v < INT_MIN)
        {
            c->bad = self;
            return POPT_ERROR_OVERFLOW;
        }
        if (o->arg)
            *(int *)o->arg = (int)v;
    }
    return o->val;
}

static inline int poptGetNextOpt(poptContext c)
{
    const struct poptOption *o;
    if (c->cluster && *c->cluster)
    {
        char s = *c->cluster++;
        o = smvp_stub_popt_find(c, NULL, 0, s);
        if (!o)
        {
            c->bad = c->argv[c->next - 1];
            return POPT_ERROR_BADOPT;
        }
        const char *rest = (*c->cluster && o->argInfo != POPT_ARG_NONE) ? c->cluster : NULL;
        if (rest)
            c->cluster = NULL;
        return smvp_stub_popt_store(c, o, rest, c->argv[c->next - 1]);
    }
    c->cluster = NULL;
    if (c->optsDone || c->next >= c->argc)
        return -1;
    const char *a = c->argv[c->next];
    if (a[0] != '-' || a[1] == '\0')
    {
        c->optsDone = 1; /* POSIXMEHARDER: first positional ends option parsing */
        return -1;
    }
    c->next++;
    if (a[1] == '-')
    {
        if (a[2] == '\0')
        {
            c->optsDone = 1;
            return -1;
        }
        const char *eq = strchr(a + 2, '=');
        size_t llen = eq ? (size_t)(eq - (a + 2)) : strlen(a + 2);
        o = smvp_stub_popt_find(c, a + 2, llen, 0);
        if (!o)
        {
            c->bad = a;
            return POPT_ERROR_BADOPT;
        }
        return smvp_stub_popt_store(c, o, eq ? eq + 1 : NULL, a);
    }
    c->cluster = a + 1;
    return poptGetNextOpt(c);
}

static inline const char *poptPeekArg(poptContext c) { return c->next < c->argc ? c->argv[c->next] : NULL; }
static inline const char *poptGetArg(poptContext c) { return c->next < c->argc ? c->argv[c->next++] : NULL; }
static inline poptContext poptFreeContext(poptContext c)
{
    free(c);
    return NULL;
}
static inline const char *poptBadOption(poptContext c, unsigned int flags)
{
    (void)flags;
    return c->bad ? c->bad : "?";
}
static inline const char *poptStrerror(int e)
{
    switch (e)
    {
    case POPT_ERROR_NOARG:
        return "missing argument";
    case POPT_ERROR_BADOPT:
        return "unknown option";
    case POPT_ERROR_BADNUMBER:
        return "invalid numeric value";
    case POPT_ERROR_OVERFLOW:
        return "number too large or too small";
    default:
        return "unknown error";
    }
}

#endif
