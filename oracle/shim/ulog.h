/*
 * Logging shim so the reference sources compile without Parrot's libulog.
 * TEST INFRASTRUCTURE ONLY (oracle build). Every macro is a no-op except the
 * two "return if" guards, which keep the reference's error-return behaviour.
 */
#ifndef ORACLE_SHIM_ULOG_H
#define ORACLE_SHIM_ULOG_H
#define ULOG_DECLARE_TAG(x)
#define ULOGD(...) do {} while (0)
#define ULOGI(...) do {} while (0)
#define ULOGN(...) do {} while (0)
#define ULOGW(...) do {} while (0)
#define ULOGE(...) do {} while (0)
#define ULOGC(...) do {} while (0)
#define ULOG_ERRNO(...) do {} while (0)
#define ULOG_ERRNO_RETURN_IF(cond, err) do { if (cond) return; } while (0)
#define ULOG_ERRNO_RETURN_ERR_IF(cond, err) do { if (cond) return -(err); } while (0)
#define ULOG_ERRNO_RETURN_VAL_IF(cond, err, val) do { if (cond) return (val); } while (0)
#endif
